// Links the C ABI of include/spl.h.  cudart is linked statically inside the library.
fn main() {
    let dir = std::env::var("SPL_LIB_DIR").expect("set SPL_LIB_DIR to the directory of libspalinalg_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=spalinalg_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
}
