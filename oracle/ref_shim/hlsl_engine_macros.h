/* TEST INFRASTRUCTURE. Pre-included into the C-preprocessor pass over a reference shader (oracle/ref_build_shaders.sh): NRD.hlsli:126-139
   lets a "custom engine" supply the resource-declaration macros; these expand to the C++ objects of hlsl_cpu.h. */
#define NRD_INTERNAL
#define NRD_CS_MAIN hlsl_main
#define NRD_CONSTANTS_START( name )
#define NRD_CONSTANT( constantType, constantName ) static constantType constantName; static hlsl::ConstRegistrar _creg_##constantName( &g_module, &constantName );
#define NRD_CONSTANTS_END
#define NRD_INPUTS_START
#define NRD_INPUT( resourceType, dataType, resourceName, regName, bindingIndex ) static resourceType<dataType> resourceName; static hlsl::TexRegistrar _treg_##resourceName( &g_module, &resourceName.t, #regName[0], bindingIndex );
#define NRD_INPUTS_END
#define NRD_OUTPUTS_START
#define NRD_OUTPUT( resourceType, dataType, resourceName, regName, bindingIndex ) static resourceType<dataType> resourceName; static hlsl::TexRegistrar _treg_##resourceName( &g_module, &resourceName.t, #regName[0], bindingIndex );
#define NRD_OUTPUTS_END
#define NRD_SAMPLERS_START
#define NRD_SAMPLER( resourceType, resourceName, regName, bindingIndex ) static resourceType resourceName = resourceType( #resourceName );
#define NRD_SAMPLERS_END
#define unorm
#define snorm
