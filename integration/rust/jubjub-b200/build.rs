// UNCOMPILED reference text (no Rust toolchain in this image).
// JUBJUB_B200_LIB_DIR points at the directory that holds libjubjub_b200.so (this repository's jubjub_b200/).
fn main() {
    let dir = std::env::var("JUBJUB_B200_LIB_DIR").expect("set JUBJUB_B200_LIB_DIR to <repo>/jubjub_b200");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=jubjub_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=JUBJUB_B200_LIB_DIR");
}
