#!/bin/bash
python tools/exp/fp32_traj.py 2>&1 | tail -12
