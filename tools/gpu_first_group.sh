#!/bin/bash
for f in 2 3 4 5 6; do echo "== first group $f windows"; P2B_MSM_FIRST_GROUP=$f P2B_ACC_VARIANT=2 python tools/msm_variants.py child 20 21 22 2>&1 | grep variant; done
