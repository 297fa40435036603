#!/bin/bash
timeout 300 python tools/quick_bench.py 2>&1 | tail -40
