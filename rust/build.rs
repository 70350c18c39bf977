// build.rs for matter-labs/hodor with `--features cuda`: links the prebuilt libhodor_b200.so
// (make -C hodor_b200/csrc in the hodor_b200 repository).  HODOR_B200_LIB_DIR must point at the
// directory that holds the library; it is also added to the binary's rpath so `cargo test` finds it.
fn main() {
    println!("cargo:rerun-if-env-changed=HODOR_B200_LIB_DIR");
    if std::env::var("CARGO_FEATURE_CUDA").is_ok() {
        let dir = std::env::var("HODOR_B200_LIB_DIR")
            .expect("set HODOR_B200_LIB_DIR to the directory containing libhodor_b200.so");
        println!("cargo:rustc-link-search=native={}", dir);
        println!("cargo:rustc-link-lib=dylib=hodor_b200");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    }
}
