// TEST INFRASTRUCTURE — host build of the product header include/nrd_frontend.cuh, evaluated through the product's own verification sequence
// (nrd_sample_b200/csrc/frontend_probe.inl) so that tests/test_frontend_codecs.py can compare it bit for bit with the reference's NRD.hlsli
// (oracle/_ref/libnrd_refshaders.so : NRD_FrontEndProbe.cs.hlsl). in: 6 arrays of n float4, out: 21 arrays of n float4.
#include "../include/nrd_frontend.cuh"
#include "../nrd_sample_b200/csrc/frontend_probe.inl"

extern "C" __attribute__((visibility("default"))) void nrd_oracle_frontend_probe(const float* const* in, float* const* out, int n) {
    for (int i = 0; i < n; i++) {
        nrdfe::F4 a[6], r[21];
        for (int k = 0; k < 6; k++) a[k] = nrdfe::f4(in[k][4 * i], in[k][4 * i + 1], in[k][4 * i + 2], in[k][4 * i + 3]);
        frontEndProbeColumn(a, r);
        for (int k = 0; k < 21; k++) {
            out[k][4 * i] = r[k].x; out[k][4 * i + 1] = r[k].y; out[k][4 * i + 2] = r[k].z; out[k][4 * i + 3] = r[k].w;
        }
    }
}
