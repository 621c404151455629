// Links libmyzkp_b200.so.  Either point MYZKP_B200_LIB_DIR at a prebuilt library or let this
// script run the in-tree Makefile (nvcc -gencode arch=compute_100a,code=sm_100a).
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    println!("cargo:rerun-if-env-changed=MYZKP_B200_LIB_DIR");
    let lib_dir = match env::var("MYZKP_B200_LIB_DIR") {
        Ok(d) => PathBuf::from(d),
        Err(_) => {
            let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
            let csrc = root.join("myzkp_b200/csrc");
            let status = Command::new("make").arg("-C").arg(&csrc).arg("-j8").status().expect("make failed to start");
            assert!(status.success(), "building libmyzkp_b200.so failed (needs nvcc 12.9+, sm_100a)");
            root.join("myzkp_b200")
        }
    };
    println!("cargo:rustc-link-search=native={}", lib_dir.display());
    println!("cargo:rustc-link-lib=dylib=myzkp_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", lib_dir.display());
}
