// build.rs for the patched `rtbvh` crate: link against this repository's librtbvh_rs.so (the CUDA library that
// exports the rtbvh_ffi C ABI plus the rtbvh_gpu_* batch extension).
//
//   RTBVH_B200_LIB_DIR=/path/to/repo/rtbvh_b200 cargo test
//
// Not compiled here (no cargo in the image); see rust/README.md.
use std::env;
use std::path::PathBuf;

fn main() {
    println!("cargo:rerun-if-env-changed=RTBVH_B200_LIB_DIR");
    let dir = env::var_os("RTBVH_B200_LIB_DIR")
        .map(PathBuf::from)
        .expect("set RTBVH_B200_LIB_DIR to the directory that holds librtbvh_rs.so (the repo's rtbvh_b200/)");
    assert!(
        dir.join("librtbvh_rs.so").exists(),
        "{} holds no librtbvh_rs.so: run `python -c 'import __graft_entry__ as g; g.build()'` in the repo first",
        dir.display()
    );
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=rtbvh_rs");
    // so that `cargo test` / `cargo run --example` find the library without LD_LIBRARY_PATH
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
}
