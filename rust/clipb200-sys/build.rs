// Links libclipb200.so (built by `make -C clip_embedder_rs_b200/csrc`).  CLIPB200_LIB_DIR overrides the search path.
fn main() {
    let dir = std::env::var("CLIPB200_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{manifest}/../../clip_embedder_rs_b200")
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=clipb200");
    println!("cargo:rerun-if-env-changed=CLIPB200_LIB_DIR");
}
