// TEST INFRASTRUCTURE. A probe "shader" of OURS that calls the application-side functions of the reference's NRD.hlsli ( front end packers,
// back end unpackers, SG / SH resolve ) on arbitrary inputs: compiled by oracle/ref_build_shaders.py like the reference's own shaders, it is
// the reference arithmetic that include/nrd_frontend.cuh is pinned against ( tests/test_frontend_codecs.py ). One thread per input column.
#include "NRD.hlsli"

NRD_INPUTS_START
    NRD_INPUT( Texture2D, float4, gIn_A, t, 0 ) // N.xyz ( any length ), roughness
    NRD_INPUT( Texture2D, float4, gIn_B, t, 1 ) // V.xyz, materialID / 3
    NRD_INPUT( Texture2D, float4, gIn_C, t, 2 ) // UNORM texel of IN_NORMAL_ROUGHNESS
    NRD_INPUT( Texture2D, float4, gIn_D, t, 3 ) // radiance.xyz ( may hold NaN / INF / negatives ), hit distance
    NRD_INPUT( Texture2D, float4, gIn_E, t, 4 ) // direction.xyz, viewZ
    NRD_INPUT( Texture2D, float4, gIn_F, t, 5 ) // misc scalars in 0..1
NRD_INPUTS_END

NRD_OUTPUTS_START
    NRD_OUTPUT( RWTexture2D, float4, gOut_0, u, 0 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_1, u, 1 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_2, u, 2 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_3, u, 3 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_4, u, 4 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_5, u, 5 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_6, u, 6 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_7, u, 7 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_8, u, 8 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_9, u, 9 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_10, u, 10 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_11, u, 11 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_12, u, 12 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_13, u, 13 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_14, u, 14 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_15, u, 15 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_16, u, 16 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_17, u, 17 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_18, u, 18 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_19, u, 19 )
    NRD_OUTPUT( RWTexture2D, float4, gOut_20, u, 20 )
NRD_OUTPUTS_END

[numthreads( 16, 1, 1 )]
void hlsl_main( int2 threadPos : SV_GroupThreadID, uint2 groupPos : SV_GroupID, int2 pixelPos : SV_DispatchThreadID, uint threadIndex : SV_GroupIndex )
{
    float4 inA = gIn_A[ pixelPos ], inB = gIn_B[ pixelPos ], inC = gIn_C[ pixelPos ], inD = gIn_D[ pixelPos ], inE = gIn_E[ pixelPos ], inF = gIn_F[ pixelPos ];

    float3 N = normalize( inA.xyz );
    float3 V = normalize( inB.xyz );
    float roughness = inA.w;
    float3 hitDistParams = float3( 3.0, 0.1, 20.0 );
    float viewZ = inE.w;
    float3 dir = normalize( inE.xyz );

    // G-buffer
    gOut_0[ pixelPos ] = NRD_FrontEnd_PackNormalAndRoughness( inA.xyz, roughness, inB.w * 3.0 );
    float materialID;
    float4 nr = NRD_FrontEnd_UnpackNormalAndRoughness( inC, materialID );
    gOut_1[ pixelPos ] = nr;
    gOut_2[ pixelPos ] = float4( REBLUR_FrontEnd_GetNormHitDist( inD.w, viewZ, hitDistParams, roughness ), _REBLUR_GetHitDistanceNormalization( viewZ, hitDistParams, roughness ),
                                 materialID, REBLUR_GetHitDist( inF.x, viewZ, hitDistParams, roughness ) );

    // REBLUR
    gOut_3[ pixelPos ] = REBLUR_FrontEnd_PackRadianceAndNormHitDist( inD.xyz, inF.y * 1.5 - 0.25, true );
    float4 sh1;
    gOut_4[ pixelPos ] = REBLUR_FrontEnd_PackSh( inD.xyz, inF.y, inE.xyz, sh1, true );
    gOut_5[ pixelPos ] = sh1;
    gOut_6[ pixelPos ] = REBLUR_FrontEnd_PackDirectionalOcclusion( inE.xyz, inF.z, true );
    gOut_7[ pixelPos ] = REBLUR_BackEnd_UnpackRadianceAndNormHitDist( float4( abs( inC.xyz ) * 4.0 - 0.5, inF.w ) );

    // RELAX
    gOut_8[ pixelPos ] = RELAX_FrontEnd_PackSh( inD.xyz, inD.w, inE.xyz, sh1, true );
    gOut_9[ pixelPos ] = sh1;

    // SIGMA
    float distanceToOccluder = inF.x > 0.9 ? NRD_FP16_MAX : inD.w;
    gOut_10[ pixelPos ] = float4( SIGMA_FrontEnd_PackPenumbra( distanceToOccluder, 0.0087 ), SIGMA_FrontEnd_PackPenumbra( distanceToOccluder, 50.0 * inF.y + 0.01, 0.5 ),
                                  SIGMA_BackEnd_UnpackShadow( inF.z ), NRD_GetNormalizedStrandThickness( inF.w * 0.01, viewZ * 0.001 ) );
    gOut_11[ pixelPos ] = SIGMA_FrontEnd_PackTranslucency( distanceToOccluder, inC.xyz * 1.5 - 0.25 );

    // SG / SH resolve
    float3 radiance = abs( inD.xyz ) + 0.01;
    radiance = NRD_IsValidRadiance( radiance ) ? min( radiance, 100.0 ) : 1.0;
    float4 s1;
    float4 s0 = REBLUR_FrontEnd_PackSh( radiance, inF.y, dir, s1, true );
    NRD_SG sg = REBLUR_BackEnd_UnpackSh( s0, s1.xyz );
    gOut_12[ pixelPos ] = float4( NRD_SG_ResolveDiffuse( sg, N, V, roughness ), NRD_ComputeCavityShadow( sg, N, inF.x, 0.9 + 0.1 * inF.z, inF.w ) );
    gOut_13[ pixelPos ] = float4( NRD_SG_ResolveSpecular( sg, N, V, roughness ), _NRD_GetSpecularDominantFactor( abs( dot( N, V ) ), roughness ) );
    gOut_14[ pixelPos ] = float4( NRD_SH_ResolveDiffuse( sg, N ), _NRD_GetSpecMagicCurve( roughness, 0.25 ) );
    gOut_15[ pixelPos ] = float4( NRD_SH_ResolveSpecular( sg, N, V, roughness ), _NRD_Luminance( radiance ) );

    float3 diffFactor, specFactor;
    NRD_MaterialFactors( N, V, inC.xyz, saturate( inF.xyz ) * 0.9 + 0.04, roughness, diffFactor, specFactor );
    gOut_16[ pixelPos ] = float4( diffFactor, 0.0 );
    gOut_17[ pixelPos ] = float4( specFactor, 0.0 );

    NRD_SG sg2 = RELAX_BackEnd_UnpackSh( float4( _NRD_LinearToYCoCg( radiance ), inD.w ), dir.zxy * _NRD_Luminance( radiance ) );
    float3 Ne = normalize( N + 0.1 * dir ), Nw = normalize( N - 0.1 * dir ), Nn = normalize( N + 0.1 * dir.yzx ), Ns = normalize( N - 0.1 * dir.yzx );
    float2 j = NRD_SG_ReJitter( sg, sg2, V, roughness, viewZ, viewZ * ( 1.0 + 0.02 * ( inF.x - 0.5 ) ), viewZ * 1.001, viewZ * 0.999, viewZ * ( 1.0 - 0.02 * ( inF.y - 0.5 ) ), N, Ne, Nw, Nn, Ns );
    gOut_18[ pixelPos ] = float4( j, NRD_SG_ExtractColor( sg2 ).xy );

    float acc = NRD_FrontEnd_SpecHitDistAveraging_Begin( );
    NRD_FrontEnd_SpecHitDistAveraging_Add( acc, NRD_FrontEnd_TrimHitDistance( inD.w, 0.5 ) );
    NRD_FrontEnd_SpecHitDistAveraging_Add( acc, inF.x > 0.5 ? 0.0 : inF.y * 10.0 );
    NRD_FrontEnd_SpecHitDistAveraging_End( acc );
    gOut_19[ pixelPos ] = float4( NRD_SG_ExtractDirection( sg2 ), acc );

    gOut_20[ pixelPos ] = _NRD_GetSphericalCapIntersection( N, 0.5 + 0.5 * inF.x, dir, 0.5 + 0.5 * inF.y );
}
