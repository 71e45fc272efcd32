// Stand-in for the un-vendored ShaderMake dependency (ShaderMake @18f5a344, NRD/CMakeLists.txt:125-130).
// The reference host library only needs the ShaderConstant POD when all NRD_EMBEDS_*_SHADERS are 0.
// TEST INFRASTRUCTURE ONLY: used by oracle/ref_build.sh to compile the reference host sources in place.
#pragma once
#include <array>
#include <cstdio>
#include <vector>
namespace ShaderMake {
struct ShaderConstant {
    const char* name;
    const char* value;
};
}
