#!/bin/sh
# Build libgolfer_b200.so from anywhere: tools/build.sh [--force] [-v]
cd "$(dirname "$0")/.." && python computer-vision-system-for-analyzing-golfer-action_b200/build.py "$@"
