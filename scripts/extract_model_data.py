#!/usr/bin/env python3
"""Extract published empirical-model CONSTANTS (numbers only, no code) from the reference tree.

Run once in the build container (where /root/reference is mounted); the outputs are committed
under iqtree2_b200/data/ so nothing at run time reads /root/reference.

  * LG exchangeabilities + frequencies (Le & Gascuel 2008) — PAML lower-triangle text embedded in
    model/modelprotein.cpp:287-309 ("model LG=").
  * C60 profile-mixture frequencies + class weights (Le, Gascuel & Lartillot 2008) —
    model/modelmixture.cpp:915-977 ("frequency C60piK = ...", "model C60 = POISSON+G+FMIX{...}").
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "iqtree2_b200", "data")


def extract_lg():
    src = open(os.path.join(REF, "model/modelprotein.cpp")).read()
    m = re.search(r"model LG=\s*\n(.*?);", src, re.S)
    nums = [float(x) for x in m.group(1).split()]
    assert len(nums) == 190 + 20, len(nums)
    with open(os.path.join(OUT, "LG.paml"), "w") as f:
        k = 0
        for i in range(1, 20):
            f.write(" ".join(repr(nums[k + j]) for j in range(i)) + "\n")
            k += i
        f.write(" ".join(repr(x) for x in nums[190:]) + "\n")


def extract_c60():
    src = open(os.path.join(REF, "model/modelmixture.cpp")).read()
    freqs = {}
    for m in re.finditer(r"frequency (C60pi\d+) = ([^;]*);", src):
        freqs[m.group(1)] = [float(x) for x in m.group(2).split()]
    m = re.search(r"model C60 = POISSON\+G\+FMIX\{([^}]*)\}", src)
    classes = []
    for item in m.group(1).split(","):
        name, rate, weight = item.split(":")
        classes.append({"name": name, "rate": float(rate), "weight": float(weight), "freq": freqs[name]})
    assert len(classes) == 60
    with open(os.path.join(OUT, "C60.json"), "w") as f:
        json.dump({"model": "C60", "base": "POISSON", "classes": classes}, f, indent=0)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    extract_lg()
    extract_c60()
    print("wrote", os.listdir(OUT))
