#!/usr/bin/env python3
"""Counts the Montgomery products (the algorithmic work unit of DESIGN.md §4: 1 Fp-mul = 300 multiply-adds) the kernel BODIES
execute per unit of work, by running them in the host mirror built with -DC12_COUNT_FP_MUL.  CPU only; input to the rooflines in
bench.py / DESIGN.md (SURVEY §8d asks for the kernel's own count beside the reference's 31.6 k per 4-pair product)."""
import ctypes, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
SRC = os.path.join(ROOT, "tests", "hostmirror", "mirror.cpp")
LIB = "/tmp/libhostmirror_count.so"
subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-DC12_COUNT_FP_MUL", "-o", LIB, SRC])
l = ctypes.CDLL(LIB)
l.hm_fp_mul_count.restype = ctypes.c_ulonglong
l.hm_fp_addsub_count.restype = ctypes.c_ulonglong
g = json.load(open(os.path.join(ROOT, "tests", "golden", "pairing.json")))
g1, g2 = bytes.fromhex(g["g1"]), bytes.fromhex(g["g2"])
out = ctypes.create_string_buffer(576 * 8)
res = {}
for k in (1, 2, 4):
    B = 8 // k
    for mode, name in ((0, "miller"), (1, "product")):
        l.hm_fp_mul_count(1)
        l.hm_fp_addsub_count(1)
        assert l.hm_pairing_product(g1, g2, B, k, mode, out) == 0
        res[f"{name}_k{k}"] = l.hm_fp_mul_count(1) / B
        res[f"{name}_k{k}_addsub"] = l.hm_fp_addsub_count(1) / B
res["final_exp"] = res["product_k1"] - res["miller_k1"]
print(json.dumps(res, indent=1))
