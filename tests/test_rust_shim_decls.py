"""rust/gpu.rs cannot be compiled here (no rustc / cargo), so its FFI surface is checked mechanically against the C
headers it binds: same functions, same arity, every argument / return type the FFI image of the C type, `#[repr(C)]`
structs with the C structs' fields — and `rust/apply_to_reference.py` is applied to a temp copy of the reference when
the reference tree is present.  No GPU, no oracle."""
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUST = os.path.join(ROOT, "rust")

SCALARS = {"size_t": "usize", "uint32_t": "u32", "uint64_t": "u64", "uint8_t": "u8", "int32_t": "i32", "int": "c_int",
           "float": "f32", "double": "f64", "void": "c_void", "char": "c_char", "unsigned char": "c_uchar", "bool": "bool"}
NAMED = {"RTAabb", "RTBvh", "RTMbvh", "RTBvhNode", "RTMbvhNode", "RTRay", "RTHit", "RTRayPacket4", "RTHitPacket4",
         "RTGpuScene", "RTGpuSceneExport", "RTTreeKind", "BvhType", "ResultCode", "RTIntersectCallback"}


def _strip_comments(src):
    return re.sub(r"//[^\n]*", "", re.sub(r"/\*.*?\*/", "", src, flags=re.S))


def c_type_to_rust(decl, with_name=True):
    """`const float *origins` -> `*const f32`; `void *const *dests` -> `*const *mut c_void`; `const float pos[3]` ->
    `*const f32`; `RTBvh bvh` -> `RTBvh`."""
    decl = " ".join(decl.replace("*", " * ").split())
    is_array = bool(re.search(r"\[\d*\]$", decl))
    decl = re.sub(r"\s*\[\d*\]$", "", decl)
    toks = decl.split()
    if with_name and toks[-1] != "*":
        toks = toks[:-1]                       # parameter name
    # base type = everything up to the first '*'
    k = toks.index("*") if "*" in toks else len(toks)
    base, rest = toks[:k], toks[k:]
    base_const = "const" in base
    base = " ".join(t for t in base if t != "const")
    rust = SCALARS.get(base, base if base in NAMED else None)
    assert rust is not None, f"unmapped C type {base!r} in {decl!r}"
    # each '*' optionally followed by 'const' (constness of THAT pointer, which Rust ignores for the outermost level)
    levels = []
    i = 0
    while i < len(rest):
        assert rest[i] == "*", decl
        ptr_const = i + 1 < len(rest) and rest[i + 1] == "const"
        levels.append(ptr_const)
        i += 2 if ptr_const else 1
    pointee_const = base_const
    for ptr_const in levels:
        rust = ("*const " if pointee_const else "*mut ") + rust
        pointee_const = ptr_const
    if is_array:
        rust = ("*const " if base_const else "*mut ") + rust
    return rust


def c_functions():
    out = {}
    for h in ("rtbvh.h", "rtbvh_gpu.h"):
        src = _strip_comments(open(os.path.join(ROOT, "include", h)).read())
        for m in re.finditer(r"\b(ResultCode|void|int|const char \*)\s*(\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
            ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
            if "(*" in args or name.startswith("("):
                continue
            params = [] if args in ("", "void") else [c_type_to_rust(a.strip()) for a in args.split(",")]
            rret = {"ResultCode": "ResultCode", "void": None, "int": "c_int", "const char *": "*const c_char"}[ret]
            assert name not in out, f"{name} declared twice"
            out[name] = (params, rret)
    return out


def rust_source():
    return open(os.path.join(RUST, "gpu.rs")).read()


def rust_functions():
    src = _strip_comments(rust_source())
    m = re.search(r'extern\s+"C"\s*\{(.*?)\n    \}', src, flags=re.S)
    assert m, "no extern \"C\" block in rust/gpu.rs"
    out = {}
    for f in re.finditer(r"pub fn (\w+)\s*\(([^;]*?)\)\s*(?:->\s*([^;]+?))?\s*;", m.group(1), flags=re.S):
        name, args, ret = f.group(1), f.group(2).strip(), f.group(3)
        params = []
        if args:
            for a in args.split(","):
                pname, ptype = a.split(":", 1)
                params.append(" ".join(ptype.split()))
        assert name not in out, f"{name} bound twice"
        out[name] = (params, ret.strip() if ret else None)
    return out


def test_extern_block_matches_the_headers():
    c, r = c_functions(), rust_functions()
    assert len(c) == 60
    assert sorted(c) == sorted(r), f"only in headers: {sorted(set(c) - set(r))}; only in gpu.rs: {sorted(set(r) - set(c))}"
    for name in sorted(c):
        assert c[name][1] == r[name][1], f"{name}: return type {r[name][1]} vs C {c[name][1]}"
        assert len(c[name][0]) == len(r[name][0]), f"{name}: arity {len(r[name][0])} vs C {len(c[name][0])}"
        for k, (ct, rt) in enumerate(zip(c[name][0], r[name][0])):
            assert ct == rt, f"{name}: argument {k} is {rt} in gpu.rs, the header says {ct}"


def test_library_exports_what_the_shim_links(tmp_path):
    from rtbvh_b200 import api
    L = api.lib()
    for name in rust_functions():
        assert hasattr(L, name), f"gpu.rs binds {name}, librtbvh_rs.so does not export it"


def _c_structs():
    out = {}
    for h in ("rtbvh.h", "rtbvh_gpu.h"):
        src = _strip_comments(open(os.path.join(ROOT, "include", h)).read())
        for m in re.finditer(r"typedef struct (\w+)\s*\{(.*?)\}\s*\1\s*;", src, flags=re.S):
            fields = []
            for decl in m.group(2).split(";"):
                decl = " ".join(decl.split())
                if not decl:
                    continue
                head = decl.split(",")[0]
                base = " ".join(head.replace("*", " * ").split()[:-1]) if "[" not in head else None
                for part in decl.split(","):
                    part = part.strip()
                    am = re.match(r"(?:(.*?)\s+)?(\w+)\[(\d+)\]$", part)
                    if am:
                        ty = am.group(1) or fields_base
                        fields_base = ty
                        fields.append((am.group(2), f"[{SCALARS[ty]}; {am.group(3)}]"))
                    else:
                        full = part if part.count(" ") else f"{base} {part}"
                        nm = full.replace("*", " ").split()[-1]
                        fields.append((nm, c_type_to_rust(full)))
            out[m.group(1)] = fields
    return out


def _rust_structs():
    src = _strip_comments(rust_source())
    out = {}
    for m in re.finditer(r"#\[repr\(C\)\]\s*(?:#\[[^\]]*\]\s*)*pub struct (\w+)\s*\{(.*?)\}", src, flags=re.S):
        fields = []
        for f in m.group(2).split(",\n"):
            f = " ".join(f.split()).rstrip(",")
            if not f:
                continue
            nm, ty = f.split(":", 1)
            fields.append((nm.replace("pub", "").strip(), ty.strip()))
        out[m.group(1)] = fields
    return out


def test_repr_c_structs_match_the_headers():
    c, r = _c_structs(), _rust_structs()
    alias = {"*const RTBvhNode": "*const RTBvhNode", "*const RTMbvhNode": "*const RTMbvhNode"}
    for name in ("RTBvh", "RTMbvh", "RTRay", "RTHit", "RTRayPacket4", "RTHitPacket4"):
        assert name in c and name in r, name
        cf = [(n, alias.get(t, t)) for n, t in c[name]]
        assert cf == r[name], f"{name}: gpu.rs {r[name]} vs header {cf}"
    # node types cross the boundary as the crate's own structs: their sizes are pinned by the same_size KAT
    assert "pub type RTAabb = Aabb<i32>;" in rust_source() and "pub type RTMbvhNode = MbvhNode;" in rust_source()


def test_constants_match_the_headers():
    src = rust_source()
    hdr = open(os.path.join(ROOT, "include", "rtbvh.h")).read() + open(os.path.join(ROOT, "include", "rtbvh_gpu.h")).read()
    for c_name, rust_name in (("Ok", "OK"), ("Error", "ERROR"), ("NoPrimitives", "NO_PRIMITIVES"),
                              ("InequalAabbsAndPrimitives", "INEQUAL_AABBS_AND_PRIMITIVES"), ("Nan", "NAN")):
        cv = int(re.search(rf"\b{c_name} = (\d+),", hdr).group(1))
        rv = int(re.search(rf"pub const {rust_name}: ResultCode = ResultCode\((\d+)\);", src).group(1))
        assert cv == rv, c_name
    for c_name, rust_name in (("LocallyOrderedClustered", "LOCALLY_ORDERED_CLUSTERED"), ("BinnedSAH", "BINNED_SAH"),
                              ("RT_TREE_BVH", "RT_TREE_BVH"), ("RT_TREE_MBVH", "RT_TREE_MBVH")):
        cv = int(re.search(rf"\b{c_name} = (\d+),", hdr).group(1))
        rv = int(re.search(rf"pub const {rust_name}: \w+ = (\d+);", src).group(1))
        assert cv == rv, c_name
    assert "#define RT_NO_HIT 0xFFFFFFFFu" in hdr and "pub const RT_NO_HIT: u32 = 0xFFFF_FFFF;" in src


def test_delimiters_balance():
    """A cheap stand-in for a parser: brackets balance outside strings, chars and comments in every Rust file."""
    for fn in ("gpu.rs", "build.rs", "benchmark_gpu.rs"):
        src = _strip_comments(open(os.path.join(RUST, fn)).read())
        src = re.sub(r'"(?:\\.|[^"\\])*"', '""', src)
        src = re.sub(r"'(?:\\.|[^'\\])'", "' '", src)
        stack = []
        pairs = {")": "(", "]": "[", "}": "{"}
        for ch in src:
            if ch in "([{":
                stack.append(ch)
            elif ch in ")]}":
                assert stack and stack.pop() == pairs[ch], f"{fn}: unbalanced {ch}"
        assert not stack, f"{fn}: unclosed {stack[-3:]}"


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="the reference tree is only present in the build container")
def test_apply_script_patches_the_reference(tmp_path):
    dst = tmp_path / "rtbvh"
    shutil.copytree("/root/reference", dst, ignore=shutil.ignore_patterns(".git", "target"))
    for dirpath, _, files in os.walk(dst):
        os.chmod(dirpath, 0o755)
        for f in files:
            os.chmod(os.path.join(dirpath, f), 0o644)
    res = subprocess.run([sys.executable, os.path.join(RUST, "apply_to_reference.py"), str(dst)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    bvh = (dst / "src" / "bvh.rs").read_text()
    body = bvh[bvh.index("pub fn construct_binned_sah"):bvh.index("// A BVH structure")]
    assert "BinnedSahBuilder::new" not in body and "LocallyOrderedClusteringBuilder::new" not in body
    assert body.count("crate::gpu::build_on_gpu(") == 4
    assert "SpatialSahBuilder::new" in bvh                      # untouched: spatial trees are built by the reference
    construct = bvh[bvh.index("pub fn construct(bvh: &Bvh)"):bvh.index("pub fn into_raw_indices(self) -> Vec<u32> {", bvh.index("pub fn construct(bvh: &Bvh)"))]
    assert "merge_nodes" not in construct and "collapse_on_gpu(bvh)" in construct
    lib = (dst / "src" / "lib.rs").read_text()
    assert "mod gpu;" in lib and "pub use gpu::*;" in lib
    assert 'links = "rtbvh_rs"' in (dst / "Cargo.toml").read_text()
    for rel in ("src/gpu.rs", "build.rs", "examples/benchmark_gpu.rs"):
        assert (dst / rel).exists()
    # a second application must refuse (anchors gone) instead of corrupting the tree
    again = subprocess.run([sys.executable, os.path.join(RUST, "apply_to_reference.py"), str(dst)], capture_output=True, text=True)
    assert again.returncode != 0
