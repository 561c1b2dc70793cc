#!/bin/bash
# compute-sanitizer over the C++ host shim (no Python in the sanitized process): every shipped kernel once.
#   tools/sanitize_host.sh [memcheck|racecheck|synccheck|initcheck] ...      (on a GPU box, from the repo root)
set -u
mkdir -p gpurun_out/sanitize
python - <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
from datum_b200 import synth
synth.synthetic_chain(256, 256, 1, probe=1)[: 6 * 256 * 256].tofile("gpurun_out/sanitize/l0_256.bin")
synth.synthetic_chain(192, 192, 1, probe=2)[: 6 * 192 * 192].tofile("gpurun_out/sanitize/l0_192.bin")
np.random.default_rng(3).integers(0, 2**32, 6 * 64 * 64, dtype=np.uint64).astype(np.uint32).tofile("gpurun_out/sanitize/argb_64.bin")
PY
D=tests/host/host_driver
for tool in "$@"; do
  echo "=== $tool"
  S="timeout 300 compute-sanitizer --tool $tool --error-exitcode 7"
  $S $D chain 256 256 8 gpurun_out/sanitize/l0_256.bin gpurun_out/sanitize/out_256.bin 2>&1 | grep -E "SUMMARY|Error|hazard|error" | head -5; echo "chain 256^2 x 8 (pair kernels with and without tile queues, tail kernel, pageable staging): exit ${PIPESTATUS[0]}"
  $S $D faces 64 64 5 gpurun_out/sanitize/argb_64.bin gpurun_out/sanitize/out_faces.bin 2>&1 | grep -E "SUMMARY|Error|hazard|error" | head -5; echo "six-image ingest + chain: exit ${PIPESTATUS[0]}"
  $S $D irradiance 64 64 8 8 gpurun_out/sanitize/argb_64.bin gpurun_out/sanitize/out_irr.bin 2>&1 | grep -E "SUMMARY|Error|hazard|error" | head -5; echo "SH9 projection + irradiance cube: exit ${PIPESTATUS[0]}"
  # stream memory operations (cuStreamWaitValue32) do not complete under compute-sanitizer: SANITIZE_SHARED=1 to try anyway
  [ "${SANITIZE_SHARED:-0}" = "1" ] && DATUM_IBL_DEVICES=0,0 DATUM_IBL_SPLIT_MIN_FACE=1 $S $D chain 192 192 7 gpurun_out/sanitize/l0_192.bin gpurun_out/sanitize/out_192.bin 2>&1 | grep -E "SUMMARY|Error|hazard|error" | head -5&& echo "one probe shared by two contexts (peer stores, last-CTA signal, stream waits): exit ${PIPESTATUS[0]}"
done
