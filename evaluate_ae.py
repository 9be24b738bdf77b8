#!/usr/bin/env python
"""evaluate_ae.py - same command line as the reference's evaluate_ae.py, running the B200-native hot path
(see dpf_nets_b200/entry.py).  Example without ShapeNet:
    python evaluate_ae.py generation/chair demo test 2048 2048 generating --synthetic 64 --no_checkpoint"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from dpf_nets_b200 import entry  # noqa: E402

if __name__ == '__main__':
    entry.evaluate_main()
