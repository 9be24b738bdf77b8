#!/usr/bin/env python
"""train_ae.py - same command line as the reference's train_ae.py, running the B200-native hot path
(see dpf_nets_b200/entry.py).  Example without ShapeNet:
    python train_ae.py generation/chair demo 1 0.000256 --synthetic 256"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from dpf_nets_b200 import entry  # noqa: E402

if __name__ == '__main__':
    entry.train_main(svr=False)
