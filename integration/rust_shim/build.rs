// build.rs for the reference crate when src/codec.rs and src/flac.rs are replaced by the shims in
// this directory.  GLC_B200_LIB_DIR = directory that holds libglc_b200.so
// (gapless_lossy_codec_b200/ of the B200 repo after `make -C gapless_lossy_codec_b200/csrc`).
fn main()
{
    let dir = std::env::var("GLC_B200_LIB_DIR").expect("set GLC_B200_LIB_DIR to the directory of libglc_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=glc_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=GLC_B200_LIB_DIR");
}
