// Links libmicroflow_cuda.so; set MICROFLOW_CUDA_LIB_DIR to the directory that holds it
// (microflow_rs_b200/ in this repository after `python -m microflow_rs_b200._build`).
fn main() {
    if let Ok(dir) = std::env::var("MICROFLOW_CUDA_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=microflow_cuda");
}
