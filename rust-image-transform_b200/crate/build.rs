// build.rs -- compiles the hand-written CUDA kernels for sm_100a and the C-ABI host layer into
// libimagekit_cuda.so, then links the crate against it.  Mirrors csrc/Makefile.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let csrc = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let lib = out.join("libimagekit_cuda.so");
    let sources = ["plan.cpp", "context.cpp", "api.cpp", "generic.cu", "fused.cu", "fused_conv.cu", "tile.cu", "up2.cu", "banded.cu", "banded_conv.cu", "banded8.cu", "banded8_conv.cu"];
    let mut cmd = Command::new(&nvcc);
    cmd.args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo"])
        .args(["-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-fno-fast-math"])
        .arg(format!("-I{}", csrc.join("../../include").display()))
        .arg(format!("-I{}", csrc.display()))
        .args(["-shared", "-cudart", "static", "-x", "cu", "-o"])
        .arg(&lib);
    for s in sources {
        cmd.arg(csrc.join(s));
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    let status = cmd.status().expect("nvcc not found: this crate has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=imagekit_cuda");
}
