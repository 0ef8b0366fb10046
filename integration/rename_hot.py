#!/usr/bin/env python3
"""Writes the objcopy --redefine-syms table that moves the reference's definitions of the hot bridge functions out of
the way (<mangled> -> refcpu_<mangled>), so integration/miracl_core_interface_b200.cpp can define them."""
import subprocess
import sys

HOT_PREFIXES = ("sum_of_products(", "double_multiply(", "pair_ate(", "pair_double_ate(", "pair_final_exponentiation(", "pow(")
HOT_MULTIPLY_FIRST_ARG = ("point1&", "point2&", "fp12&")   # multiply(big2&, ...) stays: it is Zp plumbing


def main(obj, out):
    mangled = subprocess.run(["nm", "--defined-only", obj], capture_output=True, text=True, check=True).stdout.split("\n")
    rows = []
    for line in mangled:
        parts = line.split()
        if len(parts) != 3 or parts[1] != "T":
            continue
        sym = parts[2]
        dem = subprocess.run(["c++filt", sym], capture_output=True, text=True, check=True).stdout.strip()
        name = dem.split("miracl_core::", 1)[1]
        hot = name.startswith(HOT_PREFIXES)
        if name.startswith("multiply("):
            first = name[len("multiply("):].split(",")[0]
            hot = any(first.endswith(a) for a in HOT_MULTIPLY_FIRST_ARG)
        if hot:
            rows.append(f"{sym} refcpu_{sym}")
    assert len(rows) == 9, rows
    with open(out, "w") as f:
        f.write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
