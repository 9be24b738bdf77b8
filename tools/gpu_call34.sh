#!/bin/bash
timeout 300 python tools/graph_probe.py 2>&1 | tail -24 | cut -c1-500
