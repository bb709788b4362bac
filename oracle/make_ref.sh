#!/bin/bash
# Recipe for the REAL reference as the CPU arm of bench.py (--impl reference / cpu_baseline, kind "reference").
#
# The reference is pure Python: "building" it is placing the UNMODIFIED files of the hot path's import closure where
# the GPU box can see them.  /root/reference does not exist there, so the files are copied - from where they lie, at
# build time only - into oracle/_ref/, which is git-ignored (never part of this repo's history) but travels with the
# gpurun snapshot like the built .so files.  Closure of the path (SURVEY.md 8c import recipe):
#   mdm_utils/model_util.py -> model/RAG.py (-> audio_enc.py, mlp_module.py; `import clip` is unused: stubbed),
#   model/cfg_sampler.py, diffusion/{gaussian_diffusion,respace,nn,losses}.py
# One tree per dataset (scripts/ = TED, scripts_beat/ = BEAT): both define the top-level packages model/diffusion.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${LS_REFERENCE_ROOT:-/root/reference}"
[ -d "$REF/scripts" ] || { echo "make_ref: $REF/scripts not found (nothing to do)"; exit 0; }
for pair in "ted:scripts" "beat:scripts_beat"; do
  name="${pair%%:*}"; tree="${pair##*:}"
  dst="$HERE/_ref/$name"
  rm -rf "$dst"
  mkdir -p "$dst/model" "$dst/diffusion" "$dst/mdm_utils"
  for f in model/RAG.py model/audio_enc.py model/mlp_module.py model/cfg_sampler.py \
           diffusion/gaussian_diffusion.py diffusion/respace.py diffusion/nn.py diffusion/losses.py \
           mdm_utils/model_util.py; do
    cp "$REF/$tree/$f" "$dst/$f"
  done
  # our own stub, not a reference file: RAG.py imports `clip` (line 5) and never uses it on this path
  printf '"""Stub: scripts/model/RAG.py:5 imports clip but the RAG sampling path never calls it."""\n' > "$dst/clip.py"
  (cd "$dst" && sha256sum model/*.py diffusion/*.py mdm_utils/*.py > SHA256SUMS)
done
echo "make_ref: reference files copied to $HERE/_ref (git-ignored)"
