#!/usr/bin/env python
"""Runs pytest in this process against the CPU mock (GB_TEST_MOCK_LIB must be set) and then prints how many cooperative launches the
mock made of each kernel named in GB_MOCK_COUNT (substrings of the launch expression, separated by ";"), one "COOP <name> <n>" line each.
Test infrastructure (tests/test_next_on_cpu_mock.py).  usage: run_counted.py <pytest arguments>"""
import ctypes
import os
import sys

import pytest

rc = pytest.main(sys.argv[1:])
lib = ctypes.CDLL(os.environ["GB_TEST_MOCK_LIB"])
lib.gb_mock_coop_launches.restype = ctypes.c_long
lib.gb_mock_coop_launches.argtypes = [ctypes.c_char_p]
for k in os.environ.get("GB_MOCK_COUNT", "dhop_col_kernel;dhop_fast_kernel;smat_kernel").split(";"):
    if not k:
        continue
    print("COOP", k.replace(" ", ""), lib.gb_mock_coop_launches(k.encode()), flush=True)
sys.exit(int(rc))
