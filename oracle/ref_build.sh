#!/usr/bin/env bash
# Builds the UNMODIFIED reference NRD host library (dispatch lists + constant buffers; no pixel math)
# from the sources where they lie under /root/reference into oracle/_ref/libnrd_ref.so.
# TEST INFRASTRUCTURE ONLY. Nothing is copied into the repo; oracle/_ref/ is git-ignored.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${NRD_REFERENCE_ROOT:-/root/reference}"
NRD="$REF/External/NRD"
ML="$REF/External/NRIFramework/External/MathLib"
OUT="$HERE/_ref"
[ -d "$NRD/Source" ] || { echo "reference not present, skipping"; exit 0; }
mkdir -p "$OUT"
# "../Shaders/NRDConfig.hlsli" is resolved relative to NRD/Resources/Version.h, so the shim dir is searched
# with -I for both spellings ("NRDConfig.hlsli" from Shaders/NRD.hlsli and "../Shaders/NRDConfig.hlsli").
g++ -std=c++17 -O2 -fPIC -shared -mssse3 -msse4.1 -w \
    -DNRD_EMBEDS_SPIRV_SHADERS=0 -DNRD_EMBEDS_DXIL_SHADERS=0 -DNRD_EMBEDS_DXBC_SHADERS=0 \
    -DSPIRV_SREG_OFFSET=0 -DSPIRV_BREG_OFFSET=2 -DSPIRV_UREG_OFFSET=3 -DSPIRV_TREG_OFFSET=20 \
    -DNRD_NORMAL_ENCODING=2 -DNRD_ROUGHNESS_ENCODING=1 \
    -DNRD_API='extern "C" __attribute__((visibility("default")))' \
    -I "$HERE/ref_shim" -I "$HERE/ref_shim/Shaders" -I "$NRD/Include" -I "$NRD/Source" -I "$NRD/Shaders" -I "$NRD/Resources" -I "$ML" \
    "$NRD/Source/InstanceImpl.cpp" "$NRD/Source/Reblur.cpp" "$NRD/Source/Relax.cpp" "$NRD/Source/Sigma.cpp" \
    "$NRD/Source/Reference.cpp" "$NRD/Source/Timer.cpp" "$NRD/Source/Wrapper.cpp" \
    -o "$OUT/libnrd_ref.so"
echo "built $OUT/libnrd_ref.so"
# MathLib spot-check library: the reference's ml.hlsli compiled as C++ behind C wrappers (oracle/ref_shim/ml_wrappers.cpp)
g++ -std=c++17 -O2 -fPIC -shared -mssse3 -msse4.1 -w -fvisibility=hidden -I "$ML" "$HERE/ref_shim/ml_wrappers.cpp" -o "$OUT/libml_ref.so"
echo "built $OUT/libml_ref.so"
