// build.rs -- compiles poolgen_b200/csrc (CUDA C++, sm_100a only) into libpoolgen_cuda.so and links it.
// POOLGEN_CUDA_SRC may point at a checkout of this repository; the default is two levels up from the crate.
use std::path::PathBuf;
use std::process::Command;

fn nvcc(args: &[String]) {
    let nvcc = std::env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string());
    let status = Command::new(&nvcc).args(args).status().expect("nvcc not found (CUDA 12.8+ is required for sm_100a)");
    assert!(status.success(), "nvcc failed: {:?}", args);
}

fn main() {
    let out = PathBuf::from(std::env::var("OUT_DIR").unwrap());
    let root = std::env::var("POOLGEN_CUDA_SRC")
        .map(PathBuf::from)
        .unwrap_or_else(|_| PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap()).join("../.."));
    let csrc = root.join("poolgen_b200/csrc");
    let inc = root.join("include");
    let common: Vec<String> = [
        "-O3", "-std=c++17", "-lineinfo", "--fmad=true", "-Xcompiler", "-fPIC",
        "-gencode", "arch=compute_100a,code=sm_100a",
    ]
    .iter()
    .map(|s| s.to_string())
    .chain([format!("-I{}", inc.display()), format!("-I{}", csrc.display())])
    .collect();
    let mut objs: Vec<String> = vec![];
    for src in ["pg_api.cu", "pg_ingest.cu", "pg_tables.cu", "pg_kinship.cu", "pg_nm.cu", "pg_comm.cu", "pg_text.cu", "pg_writer.cpp"] {
        let o = out.join(format!("{src}.o")).display().to_string();
        let mut a = common.clone();
        a.extend(["-c".to_string(), csrc.join(src).display().to_string(), "-o".to_string(), o.clone()]);
        nvcc(&a);
        objs.push(o);
        println!("cargo:rerun-if-changed={}", csrc.join(src).display());
    }
    // one translation unit per device allele count (pg_scan_inst.cu instantiates scan_kernel<A, K, W, P>)
    for a_dev in 2..=6 {
        let o = out.join(format!("pg_scan_a{a_dev}.o")).display().to_string();
        let mut a = common.clone();
        a.extend([
            format!("-DPG_SCAN_A={a_dev}"), "-c".to_string(), csrc.join("pg_scan_inst.cu").display().to_string(),
            "-o".to_string(), o.clone(),
        ]);
        nvcc(&a);
        objs.push(o);
    }
    for h in ["pg_scan.cuh", "pg_device.cuh", "pg_internal.h", "pg_kin.h", "pg_synth.h", "pg_ptable.h", "pg_scan_inst.cu"] {
        println!("cargo:rerun-if-changed={}", csrc.join(h).display());
    }
    println!("cargo:rerun-if-changed={}", inc.join("poolgen_cuda.h").display());
    let lib = out.join("libpoolgen_cuda.so").display().to_string();
    let mut link: Vec<String> =
        ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o"].iter().map(|s| s.to_string()).collect();
    link.push(lib);
    link.extend(objs);
    link.extend(["-cudart", "static", "-ldl"].iter().map(|s| s.to_string()));
    nvcc(&link);
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=poolgen_cuda");
}
