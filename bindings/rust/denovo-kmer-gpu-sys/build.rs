// Links the crate against libdkb.so.  No Rust toolchain exists in the build container of this
// repository, so this crate is kept in step with include/dkb.h by tests/test_host.py
// (test_rust_sys_crate_declares_every_symbol), not by cargo.
fn main() {
    let dir = std::env::var("DKB_LIB_DIR").expect("set DKB_LIB_DIR to the directory of libdkb.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=dkb");
    println!("cargo:rerun-if-env-changed=DKB_LIB_DIR");
}
