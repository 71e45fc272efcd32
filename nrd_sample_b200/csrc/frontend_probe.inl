// Verification sequence for include/nrd_frontend.cuh: every application-side function evaluated on one column of six float4 inputs, 21 float4
// results. The same text is compiled for the device (nrdcuFrontEndProbe, executor.cu) and for the host (oracle/frontend_probe.cpp, test
// infrastructure); tests/test_frontend_codecs.py compares both with the reference's NRD.hlsli running the same sequence
// (oracle/ref_shim/Shaders/NRD_FrontEndProbe.cs.hlsl).
NRDFE_FN void frontEndProbeColumn(const nrdfe::F4 in[6], nrdfe::F4 out[21]) {
    using namespace nrdfe;
    const F4 a = in[0], b = in[1], c = in[2], d = in[3], e = in[4], f = in[5];
    const F3 N = normalize(xyz(a)), V = normalize(xyz(b));
    const float roughness = a.w;
    const F3 hitDistParams = f3(3.0f, 0.1f, 20.0f);
    const float viewZ = e.w;
    const F3 dir = normalize(xyz(e));

    // G-buffer
    out[0] = NRD_FrontEnd_PackNormalAndRoughness(xyz(a), roughness, b.w * 3.0f);
    float materialID;
    out[1] = NRD_FrontEnd_UnpackNormalAndRoughness(c, materialID);
    out[2] = f4(REBLUR_FrontEnd_GetNormHitDist(d.w, viewZ, hitDistParams, roughness), _REBLUR_GetHitDistanceNormalization(viewZ, hitDistParams, roughness), materialID,
                REBLUR_GetHitDist(f.x, viewZ, hitDistParams, roughness));

    // REBLUR
    out[3] = REBLUR_FrontEnd_PackRadianceAndNormHitDist(xyz(d), f.y * 1.5f - 0.25f, true);
    F4 sh1;
    out[4] = REBLUR_FrontEnd_PackSh(xyz(d), f.y, xyz(e), sh1, true);
    out[5] = sh1;
    out[6] = REBLUR_FrontEnd_PackDirectionalOcclusion(xyz(e), f.z, true);
    out[7] = REBLUR_BackEnd_UnpackRadianceAndNormHitDist(f4(fabsf(c.x) * 4.0f - 0.5f, fabsf(c.y) * 4.0f - 0.5f, fabsf(c.z) * 4.0f - 0.5f, f.w));

    // RELAX
    out[8] = RELAX_FrontEnd_PackSh(xyz(d), d.w, xyz(e), sh1, true);
    out[9] = sh1;

    // SIGMA
    const float distanceToOccluder = f.x > 0.9f ? NRDFE_FP16_MAX : d.w;
    out[10] = f4(SIGMA_FrontEnd_PackPenumbra(distanceToOccluder, 0.0087f), SIGMA_FrontEnd_PackPenumbra(distanceToOccluder, 50.0f * f.y + 0.01f, 0.5f), SIGMA_BackEnd_UnpackShadow(f.z),
                 NRD_GetNormalizedStrandThickness(f.w * 0.01f, viewZ * 0.001f));
    out[11] = SIGMA_FrontEnd_PackTranslucency(distanceToOccluder, f3(c.x * 1.5f - 0.25f, c.y * 1.5f - 0.25f, c.z * 1.5f - 0.25f));

    // SG / SH resolve
    F3 radiance = f3(fabsf(d.x) + 0.01f, fabsf(d.y) + 0.01f, fabsf(d.z) + 0.01f);
    radiance = NRD_IsValidRadiance(radiance) ? f3(fminf(radiance.x, 100.0f), fminf(radiance.y, 100.0f), fminf(radiance.z, 100.0f)) : f3(1.0f, 1.0f, 1.0f);
    F4 s1;
    const F4 s0 = REBLUR_FrontEnd_PackSh(radiance, f.y, dir, s1, true);
    const NRD_SG sg = REBLUR_BackEnd_UnpackSh(s0, xyz(s1));
    out[12] = f4(NRD_SG_ResolveDiffuse(sg, N, V, roughness), NRD_ComputeCavityShadow(sg, N, f.x, 0.9f + 0.1f * f.z, f.w));
    out[13] = f4(NRD_SG_ResolveSpecular(sg, N, V, roughness), _NRD_GetSpecularDominantFactor(fabsf(dot(N, V)), roughness));
    out[14] = f4(NRD_SH_ResolveDiffuse(sg, N), _NRD_GetSpecMagicCurve(roughness, 0.25f));
    out[15] = f4(NRD_SH_ResolveSpecular(sg, N, V, roughness), _NRD_Luminance(radiance));

    F3 diffFactor, specFactor;
    NRD_MaterialFactors(N, V, xyz(c), saturate(xyz(f)) * 0.9f + f3(0.04f, 0.04f, 0.04f), roughness, diffFactor, specFactor);
    out[16] = f4(diffFactor, 0.0f);
    out[17] = f4(specFactor, 0.0f);

    const NRD_SG sg2 = RELAX_BackEnd_UnpackSh(f4(_NRD_LinearToYCoCg(radiance), d.w), f3(dir.z, dir.x, dir.y) * _NRD_Luminance(radiance));
    const F3 dyzx = f3(dir.y, dir.z, dir.x);
    const F3 Ne = normalize(N + 0.1f * dir), Nw = normalize(N - 0.1f * dir), Nn = normalize(N + 0.1f * dyzx), Ns = normalize(N - 0.1f * dyzx);
    const F2 j = NRD_SG_ReJitter(sg, sg2, V, roughness, viewZ, viewZ * (1.0f + 0.02f * (f.x - 0.5f)), viewZ * 1.001f, viewZ * 0.999f, viewZ * (1.0f - 0.02f * (f.y - 0.5f)), N, Ne, Nw, Nn, Ns);
    const F3 col = NRD_SG_ExtractColor(sg2);
    out[18] = f4(j.x, j.y, col.x, col.y);

    float acc = NRD_FrontEnd_SpecHitDistAveraging_Begin();
    NRD_FrontEnd_SpecHitDistAveraging_Add(acc, NRD_FrontEnd_TrimHitDistance(d.w, 0.5f));
    NRD_FrontEnd_SpecHitDistAveraging_Add(acc, f.x > 0.5f ? 0.0f : f.y * 10.0f);
    NRD_FrontEnd_SpecHitDistAveraging_End(acc);
    out[19] = f4(NRD_SG_ExtractDirection(sg2), acc);

    out[20] = _NRD_GetSphericalCapIntersection(N, 0.5f + 0.5f * f.x, dir, 0.5f + 0.5f * f.y);
}
