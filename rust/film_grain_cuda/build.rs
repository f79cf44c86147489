// Link against libfg_b200.so.  FG_B200_LIB_DIR = the directory that holds it (film_grain_b200/ of the engine
// repository after `python film_grain_b200/build.py`); the rpath makes the binary find it at run time.
fn main() {
    println!("cargo:rerun-if-env-changed=FG_B200_LIB_DIR");
    if let Ok(dir) = std::env::var("FG_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=fg_b200");
}
