// Points the linker at libdcdf_cuda.so (built by dcdf_b200/csrc/build.sh).  DCDF_CUDA_LIB_DIR = the directory holding it.
fn main() {
    if let Ok(dir) = std::env::var("DCDF_CUDA_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
    println!("cargo:rerun-if-env-changed=DCDF_CUDA_LIB_DIR");
}
