// Counterpart of crates/cdivsufsort/build.rs:1-29: instead of four C files, compile the six
// CUDA translation units of libgsa for sm_100a through the cc crate.
fn main() {
    let csrc = "../../stringsearch_b200/csrc";
    let mut build = cc::Build::new();
    build
        .cuda(true)
        .cudart("static")
        .flag("-gencode")
        .flag("arch=compute_100a,code=sm_100a")
        .flag("-std=c++17")
        .flag("-O3")
        .flag("-lineinfo") // no -rdc: every kernel lives in the translation unit that launches it, so no device-link step is needed
        .include("../../include")
        .warnings(false);
    for f in &["api.cu", "sa_build.cu", "search.cu", "verify.cu", "bwt.cu", "lcp.cu"] {
        build.file(format!("{}/{}", csrc, f));
        println!("cargo:rerun-if-changed={}/{}", csrc, f);
    }
    build.compile("libgsa.a");
    println!("cargo:rustc-link-lib=stdc++");
}
