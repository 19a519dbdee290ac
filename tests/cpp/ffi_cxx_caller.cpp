// ffi_cxx_caller.cpp — a caller written against the reference's GENERATED C++ header (rtbvh_ffi/build.rs:33-45: cbindgen
// Language::Cxx, namespace rtbvh, include guard RTBVH_HPP): every FFI name is spelled rtbvh::..., enums as
// rtbvh::ResultCode::Ok / rtbvh::BvhType::BinnedSAH.  It must compile unchanged against include/rtbvh.hpp and behave like
// tests/c/ffi_consumer.c: without a CUDA device the builders refuse (no CPU fallback); with one, create / collapse / free.
#include <cstdio>

#include "rtbvh.hpp"

#ifndef RTBVH_HPP
#error "the reference's include guard must be defined"
#endif

static bool never_stop(uint32_t, float*, void*) { return false; }

int main() {
    alignas(16) float centers[4][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}};
    rtbvh::RTBvh bvh{UINT32_MAX, 0, nullptr, 0, nullptr};
    rtbvh::RTMbvh mbvh{UINT32_MAX, 0, nullptr, 0, nullptr};
    static_assert(sizeof(rtbvh::RTAabb) == 32 && sizeof(rtbvh::RTBvhNode) == 32 && sizeof(rtbvh::RTMbvhNode) == 128, "same_size");
    if (rtbvh::create_bvh(nullptr, 0, &centers[0][0], 16, 1, rtbvh::BvhType::BinnedSAH, &bvh) != rtbvh::ResultCode::NoPrimitives) return 2;
    if (rtbvh::create_bvh(nullptr, 4, nullptr, 16, 1, rtbvh::BvhType::LocallyOrderedClustered, &bvh) != rtbvh::ResultCode::Error) return 3;
    const rtbvh::ResultCode rc = rtbvh::create_bvh(nullptr, 4, &centers[0][0], 16, 1, rtbvh::BvhType::BinnedSAH, &bvh);
    if (rtbvh_gpu_device_count() == 0) {
        if (rc != rtbvh::ResultCode::Error) return 4;
        std::printf("ok: no CUDA device, create_bvh refused\n");
        return 0;
    }
    if (rc != rtbvh::ResultCode::Ok || bvh.index_count != 4) return 5;
    if (rtbvh::create_mbvh(bvh, &mbvh) != rtbvh::ResultCode::Ok) return 6;
    const float o[3] = {0.25f, 0.25f, -1.f}, d[3] = {0, 0, 1};
    float t = 1e30f;
    if (rtbvh::intersect(bvh, o, d, &t, nullptr, never_stop) != rtbvh::ResultCode::Ok) return 7;
    if (rtbvh::intersect_mbvh(mbvh, o, d, &t, nullptr, never_stop) != rtbvh::ResultCode::Ok) return 8;
    rtbvh::free_bvh(bvh);
    rtbvh::free_mbvh(mbvh);
    std::printf("ok: create / collapse / walk / free through the rtbvh:: spellings\n");
    return 0;
}
