#!/bin/bash
# Stages a scratch copy of the reference's scripts + shipped OpenFOAM case under _ref_scratch/ (git-ignored, NOT
# gpurun-ignored) so that a gpurun call can execute the reference's own train.py / inference.py on the B200 box, where
# /root/reference does not exist.  Never committed.
set -e
cd "$(dirname "$0")/.."
rm -rf _ref_scratch && mkdir -p _ref_scratch
cp /root/reference/*.py _ref_scratch/
cp -r /root/reference/OpenFOAM-data _ref_scratch/
chmod -R u+w _ref_scratch
rm -f _ref_scratch/OpenFOAM-data/*.png
du -sh _ref_scratch
