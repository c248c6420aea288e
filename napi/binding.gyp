{
  "targets": [{
    "target_name": "pil2gpu_addon",
    "sources": ["pil2gpu_addon.cc"],
    "include_dirs": ["../include"],
    "libraries": ["-L<(module_root_dir)/../pil2_stark_js_b200", "-lpil2gpu", "-Wl,-rpath,<(module_root_dir)/../pil2_stark_js_b200"],
    "cflags_cc": ["-std=c++17", "-O2"]
  }]
}
