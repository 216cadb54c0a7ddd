// Links the prebuilt CUDA library (built by `make -C keaki_b200/csrc`: nvcc -gencode arch=compute_100a,code=sm_100a).
// KEAKI_B200_LIB_DIR overrides the search path; the default is the in-tree location relative to this crate.
fn main() {
    let dir = std::env::var("KEAKI_B200_LIB_DIR").unwrap_or_else(|_| {
        let here = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{here}/../../keaki_b200/lib")
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=keaki_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=KEAKI_B200_LIB_DIR");
}
