#!/bin/bash
timeout 600 python scripts/stress_determinism.py 30 2>&1 | tail -40
