#!/usr/bin/env python
"""Turns a checkout of meirbon/rtbvh into the GPU-backed crate.

    python rust/apply_to_reference.py /path/to/rtbvh-checkout

What it changes (nothing else is touched; the CPU builders stay in the tree, unused by `Builder`):
  * copies gpu.rs -> src/gpu.rs, build.rs -> build.rs, benchmark_gpu.rs -> examples/benchmark_gpu.rs
  * src/lib.rs: `mod gpu;` + `pub use gpu::*;`
  * Cargo.toml: `build = "build.rs"`, `links = "rtbvh_rs"`
  * src/bvh.rs: the four `...Builder::new(..).build()` calls inside `construct_binned_sah` (`:87-111`) and
    `construct_locally_ordered_clustered` (`:113-137`) become `crate::gpu::build_on_gpu(..)` (validation and
    `BuildError`s above them stay as they are), and the body of `Mbvh::construct` (`:381-404`) becomes
    `crate::gpu::collapse_on_gpu(bvh)`.  `construct_spatial_sah` keeps its CPU builder: spatial-split trees are
    built by the reference and uploaded unchanged (`GpuScene::new`).
Every edit is anchored on one line of the reference and fails loudly if the anchor is missing or ambiguous.
Written for rtbvh 0.6.2; NOT run against cargo here (the image has no Rust toolchain).
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def replace_once(text, old, new, what):
    if text.count(old) != 1:
        raise SystemExit(f"anchor for {what!r} found {text.count(old)} times (expected 1): {old!r}")
    return text.replace(old, new)


def replace_n(text, old, new, n, what):
    if text.count(old) != n:
        raise SystemExit(f"anchor for {what!r} found {text.count(old)} times (expected {n}): {old!r}")
    return text.replace(old, new)


def patch_bvh(src):
    src = replace_n(src, "Ok(BinnedSahBuilder::new(&aabbs, self.primitives, self.primitives_per_leaf).build())",
                    "crate::gpu::build_on_gpu(&aabbs, self.primitives, self.primitives_per_leaf, BuildType::BinnedSAH)",
                    2, "construct_binned_sah")
    src = replace_once(src, "Ok(LocallyOrderedClusteringBuilder::new(aabbs, self.primitives).build())",
                       "crate::gpu::build_on_gpu(aabbs, self.primitives, None, BuildType::LocallyOrderedClustered)",
                       "construct_locally_ordered_clustered (caller's aabbs)")
    src = replace_once(src, "Ok(LocallyOrderedClusteringBuilder::new(&aabbs, self.primitives).build())",
                       "crate::gpu::build_on_gpu(&aabbs, self.primitives, None, BuildType::LocallyOrderedClustered)",
                       "construct_locally_ordered_clustered (gathered aabbs)")
    head = "    pub fn construct(bvh: &Bvh) -> Self {\n"
    tail = "    pub fn into_raw_indices(self) -> Vec<u32> {"
    i = src.index(head)
    j = src.index(tail, i)
    return src[:i] + head + "        crate::gpu::collapse_on_gpu(bvh)\n    }\n\n" + src[j:]


def patch_lib(src):
    src = replace_once(src, "mod bvh_node;\n", "mod bvh_node;\nmod gpu;\n", "mod list")
    return replace_once(src, "pub use bvh_node::*;\n", "pub use bvh_node::*;\npub use gpu::*;\n", "re-exports")


def patch_cargo(src):
    return replace_once(src, 'edition = "2018"\n', 'edition = "2018"\nbuild = "build.rs"\nlinks = "rtbvh_rs"\n', "package table")


def main():
    if len(sys.argv) != 2:
        raise SystemExit(__doc__)
    root = sys.argv[1]
    for rel, fn in (("src/bvh.rs", patch_bvh), ("src/lib.rs", patch_lib), ("Cargo.toml", patch_cargo)):
        path = os.path.join(root, rel)
        with open(path) as f:
            text = f.read()
        with open(path, "w") as f:
            f.write(fn(text))
        print("patched", rel)
    for name, rel in (("gpu.rs", "src/gpu.rs"), ("build.rs", "build.rs"), ("benchmark_gpu.rs", "examples/benchmark_gpu.rs")):
        shutil.copyfile(os.path.join(HERE, name), os.path.join(root, rel))
        print("copied ", rel)


if __name__ == "__main__":
    main()
