// Links the two libraries of the repo: libfugu_gpu.so (device ABI, include/fugu_gpu.h) and libfugu_host.so (host layer,
// include/fugu_host.h: plain C++, no CUDA code; it binds the device library with dlopen on first use, so a build with
// the `host-only` feature links and maps no CUDA code at all). FUGU_GPU_LIB_DIR points at the directory holding them
// (the repo's fugu_b200/ after `make`).
fn main() {
    if let Ok(dir) = std::env::var("FUGU_GPU_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=fugu_host");
    if std::env::var("CARGO_FEATURE_HOST_ONLY").is_err() {
        println!("cargo:rustc-link-lib=dylib=fugu_gpu");
    }
    println!("cargo:rerun-if-env-changed=FUGU_GPU_LIB_DIR");
}
