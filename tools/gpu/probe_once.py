"""The five integer-pipe probes once each (for ncu): python tools/gpu/probe_once.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from crypto12381_b200 import device as dv
for i, k in enumerate(["imad", "madc_pairs", "imad_wide", "fp_mul", "fp_sqr"]):
    dv.probe(i, 200)
    print(k, dv.probe(i, 4000))
