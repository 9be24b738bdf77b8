#!/bin/bash
timeout 300 python tools/p2_probe.py 2>&1 | tail -20
