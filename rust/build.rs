// build.rs — link directives for the `b200` feature.  NOT built here (no Rust toolchain in this image).
fn main() {
    if std::env::var_os("CARGO_FEATURE_B200").is_some() {
        // libicp_b200.so built by `make -C icp_rust_b200/csrc` (nvcc, sm_100a only)
        let dir = std::env::var("ICP_B200_LIB_DIR")
            .expect("set ICP_B200_LIB_DIR to the directory holding libicp_b200.so");
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=icp_b200");
        println!("cargo:rerun-if-env-changed=ICP_B200_LIB_DIR");
    }
}
