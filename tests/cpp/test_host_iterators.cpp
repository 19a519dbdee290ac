// test_host_iterators.cpp — the HOST side of include/rtbvh.hpp, run on the CPU (no GPU call is made): the `&T` iterators
// of src/iter.rs (traverse_iter / traverse_iter_packet on Bvh and Mbvh over from_raw trees) driving
// SpatialTriangle::intersect / intersect4 exactly like examples/benchmark.rs:25-31 and :55-61.  The trees, triangles and
// rays come from files written by tests/test_cpp_host_iterators.py (trees built by the CPU oracle); the program writes
// ray.t / packet.t and the primitive each loop ends with, and the Python side compares them bit for bit with the
// oracle's own walk.
//
//   test_host_iterators <dir>     reads  <dir>/{bvh_nodes,mbvh_nodes,indices,tris,rays,packets}.bin
//                                 writes <dir>/out_{bvh,mbvh}_{single,packet}.bin
#include <cstdio>
#include <string>
#include <vector>

#include "rtbvh.hpp"

struct Tri {
    float v[9];
    rtbvh::Vec3 vertex0() const { return {v[0], v[1], v[2]}; }
    rtbvh::Vec3 vertex1() const { return {v[3], v[4], v[5]}; }
    rtbvh::Vec3 vertex2() const { return {v[6], v[7], v[8]}; }
};

template <class T>
static std::vector<T> read_all(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path);
    std::fseek(f, 0, SEEK_END);
    const long bytes = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    std::vector<T> out((size_t)bytes / sizeof(T));
    if (bytes && std::fread(out.data(), sizeof(T), out.size(), f) != out.size()) throw std::runtime_error("short read " + path);
    std::fclose(f);
    return out;
}
template <class T>
static void write_all(const std::string& path, const std::vector<T>& v) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("cannot write " + path);
    std::fwrite(v.data(), sizeof(T), v.size(), f);
    std::fclose(f);
}

template <class Tree>
static std::vector<RTHit> single_loop(const Tree& tree, const std::vector<Tri>& tris, const std::vector<RTRay>& rays) {
    std::vector<RTHit> out(rays.size());
    for (size_t i = 0; i < rays.size(); i++) {
        rtbvh::Ray ray = rtbvh::Ray::make(rays[i].origin, rays[i].direction);
        ray.t_min = rays[i].t_min;
        ray.t = rays[i].t;
        uint32_t best = RT_NO_HIT;
        auto it = tree.traverse_iter(ray, tris.data(), tris.size());
        const Tri* tri;
        while (it.next(&tri))
            if (rtbvh::intersect(*tri, ray)) best = (uint32_t)(tri - tris.data());
        out[i] = RTHit{ray.t, best};
    }
    return out;
}

template <class Tree>
static std::vector<RTHitPacket4> packet_loop(const Tree& tree, const std::vector<Tri>& tris, const std::vector<RTRayPacket4>& packets) {
    std::vector<RTHitPacket4> out(packets.size());
    const float t_min[4] = {1e-4f, 1e-4f, 1e-4f, 1e-4f};  // Vec4::splat(1e-4), benchmark.rs:58
    for (size_t i = 0; i < packets.size(); i++) {
        const RTRayPacket4& in = packets[i];
        rtbvh::RayPacket4 p = rtbvh::make_packet(in.origin_x, in.origin_y, in.origin_z, in.direction_x, in.direction_y, in.direction_z);
        for (int l = 0; l < 4; l++) p.t[l] = in.t[l];
        RTHitPacket4 h;
        for (int l = 0; l < 4; l++) h.prim[l] = RT_NO_HIT;
        auto it = tree.traverse_iter_packet(p, tris.data(), tris.size());
        const Tri* tri;
        while (it.next(&tri)) {
            const unsigned m = rtbvh::intersect4(*tri, p, t_min);
            for (int l = 0; l < 4; l++)
                if (m & (1u << l)) h.prim[l] = (uint32_t)(tri - tris.data());
        }
        for (int l = 0; l < 4; l++) h.t[l] = p.t[l];
        out[i] = h;
    }
    return out;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string dir = argv[1];
    try {
        const auto bnodes = read_all<RTBvhNode>(dir + "/bvh_nodes.bin");
        const auto mnodes = read_all<RTMbvhNode>(dir + "/mbvh_nodes.bin");
        const auto indices = read_all<uint32_t>(dir + "/indices.bin");
        const auto tris = read_all<Tri>(dir + "/tris.bin");
        const auto rays = read_all<RTRay>(dir + "/rays.bin");
        const auto packets = read_all<RTRayPacket4>(dir + "/packets.bin");
        const rtbvh::Bvh bvh = rtbvh::Bvh::from_raw(bnodes.data(), bnodes.size(), indices.data(), indices.size(), rtbvh::BuildType::BinnedSAH);
        const rtbvh::Mbvh mbvh = rtbvh::Mbvh::from_raw(mnodes.data(), mnodes.size(), indices.data(), indices.size());
        if (!bvh.validate(tris.size())) throw std::runtime_error("validate failed");
        if (bvh.into_raw().first.size() != bnodes.size() || mbvh.into_raw().second.size() != indices.size())
            throw std::runtime_error("into_raw sizes");
        write_all(dir + "/out_bvh_single.bin", single_loop(bvh, tris, rays));
        write_all(dir + "/out_mbvh_single.bin", single_loop(mbvh, tris, rays));
        write_all(dir + "/out_bvh_packet.bin", packet_loop(bvh, tris, packets));
        write_all(dir + "/out_mbvh_packet.bin", packet_loop(mbvh, tris, packets));
        // the Bvh flavours reject an empty primitive slice (iter.rs:36-44), the Mbvh flavours do not look at it
        rtbvh::Ray r = rtbvh::Ray::make(rays[0].origin, rays[0].direction);
        const Tri* tri;
        if (bvh.traverse_iter(r, tris.data(), 0).next(&tri)) throw std::runtime_error("empty primitives must yield nothing (Bvh)");
        // NaN origin: rejected by the Bvh iterators (iter.rs:36-39)
        const float nan_o[3] = {0.f, NAN, 0.f};
        rtbvh::Ray rn = rtbvh::Ray::make(nan_o, rays[0].direction);
        if (bvh.traverse_iter(rn, tris.data(), tris.size()).next(&tri)) throw std::runtime_error("NaN rays must yield nothing (Bvh)");
    } catch (const std::exception& e) {
        std::fprintf(stderr, "FAILED: %s\n", e.what());
        return 1;
    }
    std::printf("ok\n");
    return 0;
}
