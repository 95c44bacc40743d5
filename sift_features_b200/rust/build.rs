// Links against the in-tree libsift_b200.so (built by `python -m sift_features_b200.build`).
fn main() {
    let dir = std::env::var("SIFT_B200_LIB_DIR").unwrap_or_else(|_| "..".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=sift_b200");
    println!("cargo:rerun-if-env-changed=SIFT_B200_LIB_DIR");
}
