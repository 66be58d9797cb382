// Links libphastft_cuda.so.  PHASTFT_CUDA_LIB_DIR points at the directory holding it
// (default: ../phastft_b200 relative to this crate).
fn main() {
    let dir = std::env::var("PHASTFT_CUDA_LIB_DIR").unwrap_or_else(|_| {
        let manifest = std::env::var("CARGO_MANIFEST_DIR").unwrap();
        format!("{manifest}/../phastft_b200")
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=phastft_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=PHASTFT_CUDA_LIB_DIR");
}
