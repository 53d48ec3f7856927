fn main() {
    // directory holding libckks_b200.so (built by `python -c 'import __graft_entry__ as g; g.build()'`)
    let dir = std::env::var("CKKS_B200_LIB_DIR").expect("set CKKS_B200_LIB_DIR to toy-heaan-ckks_b200/");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=ckks_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
}
