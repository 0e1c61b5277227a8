// Build script of genfer-taylor-sys.
//
// Two modes:
//  * GENFER_TAYLOR_LIB_DIR=<dir containing libgenfer_taylor.so> : link the prebuilt library
//    (what `python genfer_b200/build.py` produces in-tree).
//  * otherwise: compile genfer_b200/csrc/*.cu with nvcc through the `cc` crate for sm_100a only
//    (no multi-arch fatbin, no fallback path) and link the result statically.
use std::{env, path::PathBuf};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    println!("cargo:rerun-if-env-changed=GENFER_TAYLOR_LIB_DIR");
    if let Ok(dir) = env::var("GENFER_TAYLOR_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=genfer_taylor");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    } else {
        let csrc = root.join("genfer_b200/csrc");
        let mut b = cc::Build::new();
        b.cuda(true)
            .cudart("static")
            .flag("-gencode")
            .flag("arch=compute_100a,code=sm_100a")
            .flag("-lineinfo")
            .flag("-std=c++17")
            .opt_level(3)
            .include(root.join("include"));
        // every translation unit of the library: csrc/*.cu and csrc/*.cpp (the host evaluator behind gtp_run_sgcl)
        let mut sources: Vec<PathBuf> = std::fs::read_dir(&csrc)
            .expect("genfer_b200/csrc")
            .filter_map(|e| e.ok().map(|e| e.path()))
            .filter(|p| matches!(p.extension().and_then(|x| x.to_str()), Some("cu") | Some("cpp")))
            .collect();
        sources.sort();
        for p in sources {
            println!("cargo:rerun-if-changed={}", p.display());
            b.file(p);
        }
        b.compile("genfer_taylor");
        println!("cargo:rustc-link-lib=dylib=stdc++");
    }
    #[cfg(feature = "regen-bindings")]
    {
        let header = root.join("include/genfer_taylor.h");
        println!("cargo:rerun-if-changed={}", header.display());
        bindgen::Builder::default()
            .header(header.to_str().unwrap())
            .allowlist_function("gt[pu]_.*")
            .allowlist_type("gt[pu]_.*")
            .allowlist_var("GTP_.*")
            .generate()
            .expect("bindgen")
            .write_to_file(PathBuf::from(env::var("OUT_DIR").unwrap()).join("bindings.rs"))
            .unwrap();
    }
}
