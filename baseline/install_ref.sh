#!/bin/sh
# Installs the UNMODIFIED reference (honglu2875/hironaka, /root/reference) into baseline/_ref/ (git-ignored,
# travels to the GPU box with gpurun).  No dependency is resolved: the torch path of the reference needs
# only torch, numpy, scipy, pyyaml (present); jax / chex / gym are stubbed at import time by
# baseline/reference_arm.py exactly as SURVEY.md section 8c describes.
set -e
cd "$(dirname "$0")/.."
REF="${HIRONAKA_REFERENCE:-/root/reference}"
rm -rf baseline/_ref
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target baseline/_ref "$REF"
