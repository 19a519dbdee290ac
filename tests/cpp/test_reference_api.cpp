// The reference's own tests, written against the C++ mirror (include/rtbvh.hpp) so they read like the originals:
//   src/lib.rs:28-154, 246-311          test_invalid_input, test_sah, test_locb, five_triangle_test_case
//   src/builders/binned_sah.rs:408-458  no_primitives, test_binned_sah_build      (teapot)
//   src/builders/locb.rs:337-388        no_primitives, test_locb_build            (teapot)
//   rtbvh_ffi/src/lib.rs:856-1019       same_size, create_delete, intersect
// plus: the batched GPU traversal equals the reference's iterator loop written with the mirrored iterators.
// Usage: test_reference_api <teapot_tris.npy>     (needs a CUDA device: the builders run on the GPU)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>

#include "rtbvh.hpp"

using namespace rtbvh;

static int g_checks = 0;
#define CHECK(c)                                                              \
    do {                                                                      \
        g_checks++;                                                           \
        if (!(c)) {                                                           \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); \
            std::exit(1);                                                     \
        }                                                                     \
    } while (0)

struct Triangle {  // the tests' Triangle (src/lib.rs:159-195)
    Vec3 v0, v1, v2;
    Vec3 center() const { return (v0 + v1 + v2) * (1.0f / 3.0f); }
    Aabb aabb() const {
        Aabb bb = Aabb::empty();
        bb.grow(v0);
        bb.grow(v1);
        bb.grow(v2);
        return bb;
    }
    Vec3 vertex0() const { return v0; }
    Vec3 vertex1() const { return v1; }
    Vec3 vertex2() const { return v2; }
};

static std::vector<Triangle> load_teapot(const char* npy) {
    std::ifstream f(npy, std::ios::binary);
    CHECK(f.good());
    char magic[8];
    f.read(magic, 8);
    uint16_t hlen = 0;
    f.read(reinterpret_cast<char*>(&hlen), 2);
    f.seekg(10 + hlen);
    std::vector<Triangle> t(6320);
    f.read(reinterpret_cast<char*>(t.data()), (std::streamsize)(t.size() * sizeof(Triangle)));
    CHECK(f.gcount() == (std::streamsize)(t.size() * sizeof(Triangle)));
    return t;
}

static std::vector<Triangle> quad(float z) {
    const Vec3 v[4] = {{-1, -1, z}, {1, -1, z}, {1, 1, z}, {-1, 1, z}};
    return {Triangle{v[0], v[1], v[2]}, Triangle{v[0], v[2], v[3]}};
}

template <class T>
static Builder<T> builder(const std::vector<T>& prims, const std::vector<Aabb>* aabbs = nullptr, size_t leaf = 0) {
    Builder<T> b;
    if (aabbs) b.aabbs = std::make_pair(aabbs->data(), aabbs->size());
    b.primitives = prims.data();
    b.primitive_count = prims.size();
    b.primitives_per_leaf = leaf;
    return b;
}

static void test_invalid_input() {  // src/lib.rs:28-64
    std::vector<Triangle> none;
    CHECK(builder(none).construct_binned_sah().unwrap_err() == BuildError{BuildError::NoPrimitives});
    std::vector<Triangle> one(1);
    CHECK(builder(one).construct_binned_sah().is_ok());
    std::vector<Aabb> empty;
    CHECK((builder(one, &empty).construct_binned_sah().unwrap_err() == BuildError{BuildError::InequalAabbsAndPrimitives, 0, 1}));
}

static void test_sah_and_locb() {  // src/lib.rs:66-124
    auto prims = quad(0.0f);
    std::vector<Aabb> aabbs;
    for (auto& t : prims) aabbs.push_back(t.aabb());
    CHECK(builder(prims, &aabbs).construct_binned_sah().is_ok());
    CHECK(builder(prims, &aabbs).construct_locally_ordered_clustered().is_ok());
}

static void five_triangle_test_case() {  // src/lib.rs:246-311
    std::vector<Triangle> t = {
        {{128.79f, -1422.82f, 0.16f}, {128.5f, -1426.88f, 0.16f}, {128.79f, -1426.9067f, 0.16f}},
        {{129.8f, -1422.8629f, 0.16f}, {128.79f, -1422.82f, 0.16f}, {128.79f, -1426.9067f, 0.16f}},
        {{129.8f, -1422.8629f, 0.16f}, {128.79f, -1426.9067f, 0.16f}, {129.8f, -1427.0f, 0.16f}},
        {{130.2f, -1422.88f, 0.16f}, {129.8f, -1422.8629f, 0.16f}, {129.8f, -1427.0f, 0.16f}},
        {{130.2f, -1422.88f, 0.16f}, {129.8f, -1427.0f, 0.16f}, {130.2f, -1423.13f, 0.16f}},
    };
    for (size_t i = 1; i <= 10; i++) {
        Bvh bvh = builder(t, nullptr, i).construct_binned_sah().unwrap();
        CHECK(bvh.validate(t.size()));
        Mbvh mbvh(bvh);
        CHECK(mbvh.quad_node_count() >= 1);
    }
}

static void teapot_builds(const std::vector<Triangle>& prims) {  // binned_sah.rs:408-458, locb.rs:337-388
    std::vector<Aabb> aabbs;
    for (auto& t : prims) aabbs.push_back(t.aabb());
    for (int kind = 0; kind < 2; kind++) {
        auto b = builder(prims, &aabbs);
        Bvh bvh = (kind ? b.construct_binned_sah() : b.construct_locally_ordered_clustered()).unwrap();
        CHECK(bvh.node_count() >= aabbs.size() && bvh.node_count() <= 2 * aabbs.size());
        const Aabb bounds = bvh.bounds();
        CHECK(bounds.is_valid());
        CHECK(bvh.validate(prims.size()));
        for (auto& t : prims) CHECK(bounds.contains(t.vertex0()) && bounds.contains(t.vertex1()) && bounds.contains(t.vertex2()));
    }
}

// rtbvh_ffi/src/lib.rs:1021-1060: the FFI test's callback
struct UserData {
    Vec3 origin, direction;
    const std::vector<Triangle>* tris;
};
static bool intersect_test(uint32_t id, float* t, void* data) {
    const UserData* u = static_cast<const UserData*>(data);
    Ray r = Ray::make(&u->origin.x, &u->direction.x);
    r.t_min = 1e-5f;
    r.t = *t;
    if (intersect((*u->tris)[id], r)) *t = r.t;
    return false;
}

static void ffi_tests() {
    static_assert(sizeof(RTBvhNode) == 32 && sizeof(RTMbvhNode) == 128 && sizeof(RTAabb) == 32, "same_size");
    // create_delete (lib.rs:869-943)
    std::vector<Vec3> vertices;
    for (int x = 0; x <= 9; x++)
        for (int y = 0; y <= 9; y++) vertices.push_back(Vec3{(float)x, (float)y, 0.0f});
    std::vector<Aabb> aabbs;
    for (int i = 0; i < 27; i++) aabbs.push_back(aabb_of(vertices[i * 3], vertices[i * 3 + 1], vertices[i * 3 + 2]));
    std::vector<float> centers(27 * 4, 0.0f);
    for (int i = 0; i < 27; i++) {
        const Vec3 c = aabbs[i].center();
        centers[4 * i] = c.x; centers[4 * i + 1] = c.y; centers[4 * i + 2] = c.z;
    }
    RTBvh bvh{UINT32_MAX, 0, nullptr, 0, nullptr};
    CHECK(create_bvh(nullptr, 27, nullptr, 16, 1, BinnedSAH, &bvh) == Error);
    CHECK(create_bvh(nullptr, 27, centers.data(), 16, 1, BinnedSAH, &bvh) == Ok);
    free_bvh(bvh);
    CHECK(create_bvh(aabbs.data(), 27, centers.data(), 16, 1, BinnedSAH, &bvh) == Ok);
    RTMbvh mbvh{UINT32_MAX, 0, nullptr, 0, nullptr};
    CHECK(create_mbvh(bvh, &mbvh) == Ok);
    free_bvh(bvh);
    free_mbvh(mbvh);
    // intersect (lib.rs:946-1019)
    auto tris = quad(1.0f);
    std::vector<Aabb> qa;
    std::vector<float> qc;
    for (auto& t : tris) {
        qa.push_back(aabb_of(t.v0, t.v1, t.v2));
        const Vec3 c = qa.back().center();
        qc.insert(qc.end(), {c.x, c.y, c.z});
    }
    CHECK(create_bvh(qa.data(), 2, qc.data(), 12, 1, BinnedSAH, &bvh) == Ok);
    CHECK(create_mbvh(bvh, &mbvh) == Ok);
    UserData ud{Vec3{0, 0, 0}, Vec3{0, 0, 1}, &tris};
    float t = 1e26f;
    CHECK(intersect(bvh, &ud.origin.x, &ud.direction.x, &t, &ud, intersect_test) == Ok);
    CHECK(std::fabs(t - 1.0f) < 1.1920929e-7f);
    t = 1e26f;
    CHECK(intersect_mbvh(mbvh, &ud.origin.x, &ud.direction.x, &t, &ud, intersect_test) == Ok);
    CHECK(std::fabs(t - 1.0f) < 1.1920929e-7f);
    const float nan_o[3] = {NAN, 0, 0};
    CHECK(intersect(bvh, nan_o, &ud.direction.x, &t, &ud, intersect_test) == Nan);
    free_bvh(bvh);
    free_mbvh(mbvh);
}

// examples/benchmark.rs:25-31 written with the mirrored iterators vs the batched GPU call
static void batch_equals_iterator_loop(const std::vector<Triangle>& prims) {
    Bvh bvh = builder(prims, nullptr, 1).construct_binned_sah().unwrap();
    Mbvh mbvh(bvh);
    Scene scene(&bvh, &mbvh, &prims[0].v0.x, 12, prims.size());
    std::vector<RTRay> rays;
    uint64_t s = 12345;
    auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (float)((s >> 40) & 0xFFFFFF) / 16777216.0f; };
    for (int i = 0; i < 20000; i++) {
        const float ox = -8 + 16 * rnd(), oy = -4 + 12 * rnd(), oz = -8 + 16 * rnd();
        const float tx = -3 + 6 * rnd(), ty = 3 * rnd(), tz = -2 + 4 * rnd();
        float dx = tx - ox, dy = ty - oy, dz = tz - oz;
        const float il = 1.0f / std::sqrt(dx * dx + dy * dy + dz * dz);
        rays.push_back(RTRay{{ox, oy, oz}, 1e-4f, {dx * il, dy * il, dz * il}, 1e34f});
    }
    for (int tree = 0; tree < 2; tree++) {
        const std::vector<RTHit> hits = scene.intersect(rays, tree ? RT_TREE_MBVH : RT_TREE_BVH);
        for (size_t i = 0; i < rays.size(); i++) {
            Ray ray = Ray::make(rays[i].origin, rays[i].direction);
            uint32_t best = RT_NO_HIT, prim;
            auto body = [&](uint32_t id) {
                const float before = ray.t;
                Ray probe = ray;
                probe.t = 1e34f;
                if (intersect(prims[id], ray)) best = id;
                else if (best != RT_NO_HIT && intersect(prims[id], probe) && probe.t == before && id < best) best = id;
            };
            if (tree) {
                auto it = mbvh.traverse_iter_indices(ray);
                while (it.next(&prim)) body(prim);
            } else {
                auto it = bvh.traverse_iter_indices(ray);
                while (it.next(&prim)) body(prim);
            }
            CHECK(hits[i].t == ray.t);
            CHECK(hits[i].prim == best);
        }
    }
}

// A resident scene (built on the device, no host mirror) answers like the host-mirrored trees, through the blocking and
// the submit/wait calls, and follows moving vertices through refit like a rebuilt-and-collapsed reference tree would.
static void resident_scene_async_and_refit(const std::vector<Triangle>& prims) {
    Bvh bvh = builder(prims, nullptr, 1).construct_binned_sah().unwrap();
    Mbvh mbvh(bvh);
    Scene mirrored(&bvh, &mbvh, &prims[0].v0.x, 12, prims.size());
    Scene resident = Scene::build(&prims[0].v0.x, 12, prims.size(), BinnedSAH, 1, true);
    const std::vector<RTMbvhNode> rn = resident.read_mbvh_nodes();
    CHECK(rn.size() == mbvh.raw().node_count);
    CHECK(std::memcmp(rn.data(), mbvh.raw().nodes, rn.size() * sizeof(RTMbvhNode)) == 0);
    std::vector<RTRay> rays;
    uint64_t s = 777;
    auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (float)((s >> 40) & 0xFFFFFF) / 16777216.0f; };
    for (int i = 0; i < 50000; i++) {
        const float ox = -8 + 16 * rnd(), oy = -4 + 12 * rnd(), oz = -8 + 16 * rnd();
        const float tx = -3 + 6 * rnd(), ty = 3 * rnd(), tz = -2 + 4 * rnd();
        float dx = tx - ox, dy = ty - oy, dz = tz - oz;
        const float il = 1.0f / std::sqrt(dx * dx + dy * dy + dz * dz);
        rays.push_back(RTRay{{ox, oy, oz}, 1e-4f, {dx * il, dy * il, dz * il}, 1e34f});
    }
    const std::vector<RTHit> want = mirrored.intersect(rays);
    std::vector<RTHit> a(rays.size()), b(rays.size());
    const uint64_t t0 = resident.intersect_async(rays.data(), rays.size(), a.data());
    const uint64_t t1 = resident.intersect_async(rays.data(), rays.size(), b.data(), RT_TREE_BVH);
    CHECK(t1 > t0);
    resident.wait(t1);
    resident.wait(t0);
    CHECK(std::memcmp(a.data(), want.data(), want.size() * sizeof(RTHit)) == 0);
    const std::vector<RTHit> want_bvh = mirrored.intersect(rays, RT_TREE_BVH);
    CHECK(std::memcmp(b.data(), want_bvh.data(), want_bvh.size() * sizeof(RTHit)) == 0);
    // move every triangle, refit both scenes' worth of state: FFI refit on the host-mirrored Bvh + a new collapse vs scene refit
    std::vector<Triangle> moved = prims;
    std::vector<Aabb> boxes;
    for (size_t i = 0; i < moved.size(); i++) {
        const float d = 0.05f * std::sin((float)i * 0.37f);
        for (Vec3* v : {&moved[i].v0, &moved[i].v1, &moved[i].v2}) { v->x += d; v->y -= d * 0.5f; v->z += d * 0.25f; }
        boxes.push_back(moved[i].aabb());  // Primitive::aabb of the tests' Triangle: un-padded, what the scene refit computes
    }
    resident.refit(&moved[0].v0.x, 12, moved.size());
    bvh.refit(boxes.data());
    Mbvh mbvh2(bvh);
    Scene mirrored2(&bvh, &mbvh2, &moved[0].v0.x, 12, moved.size());
    const std::vector<RTMbvhNode> rn2 = resident.read_mbvh_nodes();
    CHECK(rn2.size() == mbvh2.raw().node_count);
    CHECK(std::memcmp(rn2.data(), mbvh2.raw().nodes, rn2.size() * sizeof(RTMbvhNode)) == 0);
    const std::vector<RTHit> h1 = resident.intersect(rays), h2 = mirrored2.intersect(rays);
    CHECK(std::memcmp(h1.data(), h2.data(), h1.size() * sizeof(RTHit)) == 0);
}

int main(int argc, char** argv) {
    CHECK(argc >= 2);
    CHECK(rtbvh_gpu_device_count() > 0);
    const auto teapot = load_teapot(argv[1]);
    test_invalid_input();
    test_sah_and_locb();
    five_triangle_test_case();
    teapot_builds(teapot);
    ffi_tests();
    batch_equals_iterator_loop(teapot);
    resident_scene_async_and_refit(teapot);
    std::printf("ok: %d checks\n", g_checks);
    return 0;
}
