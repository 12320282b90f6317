#!/bin/bash
python scripts/e2e_host_profile.py 2>&1 | tail -45
