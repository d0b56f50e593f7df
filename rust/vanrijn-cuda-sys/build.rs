// Links the prebuilt libvanrijn_cuda.so (built by `make` at the repository root with nvcc for sm_100a).
fn main() {
    let dir = std::env::var("VANRIJN_CUDA_LIB_DIR").unwrap_or_else(|_| "../../vanrijn_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=vanrijn_cuda");
    println!("cargo:rerun-if-env-changed=VANRIJN_CUDA_LIB_DIR");
}
