#!/usr/bin/env python3
"""The FP64 issue-rate microbenchmark alone (for ncu): python tools/peak_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ode_jl_b200 as m
L = m.lib()
print("DFMA instr/s      %.4g" % L.ode_b200_fp64_peak(0, 0))
print("DADD+DMUL instr/s %.4g" % L.ode_b200_fp64_peak(0, 1))
