// Links libfugu_gpu.so. FUGU_GPU_LIB_DIR points at the directory holding it (the repo's fugu_b200/ after `make`).
fn main() {
    if let Ok(dir) = std::env::var("FUGU_GPU_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=fugu_gpu");
    println!("cargo:rerun-if-env-changed=FUGU_GPU_LIB_DIR");
}
