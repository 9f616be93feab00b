#!/bin/bash
B200LU_PANEL_DBG=1 timeout 120 python scripts/prof_driver.py 8192 lu 2>&1 | grep pdbg | tail -34
