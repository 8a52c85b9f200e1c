// Links libpbrt_b200.so (built by `make -C pbrt-rs_b200`).
fn main() {
    let dir = std::env::var("PB2_LIB_DIR").unwrap_or_else(|_| "../pbrt-rs_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=pbrt_b200");
}
