#!/bin/bash
timeout 600 python tools/debug_fork_race.py 2>&1 | grep -v Warning | tail -8 | cut -c1-400
