"""TEST INFRASTRUCTURE. Turns ONE C-preprocessed reference compute shader (HLSL text on stdin or passed in) into C++ that compiles against
hlsl_cpu.h. Purely lexical: HLSL spellings that C++ lacks are rewritten, the algorithm text is untouched.

    attributes [numthreads] [unroll] ...      -> removed ( numthreads is recorded )
    parameter semantics ": SV_GroupID"         -> removed ( recorded to build the entry wrapper )
    out / inout parameters                     -> C++ references
    floating literals 0.5                      -> 0.5f ( HLSL literals are fp32 )
    scalar swizzles  s.xxx                     -> hlsl_splat3( s )
    groupshared                                -> static thread_local ( one OS thread executes a group )
    per-thread mutable statics of ml.hlsli     -> per-fiber state
    any( uv != mirrorUv )                      -> hlsl::probeMirror( ... ) ( counts the branch, value unchanged )
    uint2( GetUint( ), GetUint( ) )            -> uint2{ ... } ( left-to-right argument evaluation, as DXC does )
The output is piped into g++ by oracle/ref_build_shaders.py and never written into the repository."""
import re
import sys

SEMANTICS = {"SV_GroupThreadID": "gtid", "SV_GroupThreadId": "gtid", "SV_GroupID": "gid", "SV_GroupId": "gid", "SV_DispatchThreadID": "dtid", "SV_DispatchThreadId": "dtid",
             "SV_GroupIndex": "gi"}
FLOAT_LITERAL = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")


def transform(text: str, identifier: str, namespace: str) -> str:
    text = re.sub(r"^\s*#\s*(pragma|line).*$", "", text, flags=re.M)

    m = re.search(r"\[\s*numthreads\s*\(([^\]]*)\)\s*\]", text)
    if not m:
        raise SystemExit(f"{identifier}: no [numthreads]")
    threads = [int(eval(x, {"__builtins__": {}})) for x in m.group(1).split(",")]
    text = text[:m.start()] + text[m.end():]
    text = re.sub(r"\[\s*(unroll|loop|branch|flatten|fastopt|allow_uav_condition)\s*(\(\s*\w+\s*\))?\s*\]", "", text)

    # entry point: void hlsl_main( uint2 threadPos : SV_GroupThreadID, ... )
    m = re.search(r"void\s+hlsl_main\s*\(([^)]*)\)", text)
    if not m:
        raise SystemExit(f"{identifier}: entry point not found")
    args, params = [], []
    for p in m.group(1).split(","):
        pm = re.match(r"\s*(\w+)\s+(\w+)\s*:\s*(\w+)\s*$", p)
        if not pm:
            raise SystemExit(f"{identifier}: cannot parse entry parameter '{p}'")
        typ, name, sem = pm.groups()
        params.append(f"{typ} {name}")
        args.append(f"{typ}( {SEMANTICS[sem]} )")
    text = text[:m.start()] + "void hlsl_main( " + ", ".join(params) + " )" + text[m.end():]

    # "0.02.xxx" ( a macro constant with a swizzle ) and "name.xxx" on scalars
    text = re.sub(r"(?<![\w.])(\d+\.\d*(?:[eE][-+]?\d+)?)\s*\.\s*(x{2,4})\b", lambda k: f"hlsl_splat{len(k.group(2))}( {k.group(1)} )", text)
    text = re.sub(r"(?<![\w.\])])([A-Za-z_]\w*)\s*\.\s*(x{2,4})\b(?!\s*\()", lambda k: f"hlsl_splat{len(k.group(2))}( {k.group(1)} )", text)
    text = FLOAT_LITERAL.sub(lambda k: k.group(1) + "f", text)

    # HLSL lets a scalar answer to ".x": drop it on names that are only ever declared as scalars
    scalars = set(re.findall(r"\b(?:float|uint|int)\s+(\w+)\s*(?:=|;|,|\))", text)) - set(re.findall(r"\b(?:float|uint|int|bool)[234]\s+(\w+)", text))
    for name in scalars:
        text = re.sub(r"(?<![\w.\])])" + name + r"\s*\.\s*[xr]\b(?!\s*\()", name, text)

    # out / inout -> references
    text = re.sub(r"\b(?:inout|out)\s+((?:const\s+)?[A-Za-z_]\w*(?:\s*<[^<>]*>)?)\s+(\w+)", r"\1& \2", text)
    text = re.sub(r"([(,]\s*)in\s+((?:const\s+)?(?:float|int|uint|bool|half)\w*\s+\w+)", r"\1\2", text)

    # per-thread mutable statics ( ml.hlsli Rng ) live in the fiber
    for name in ("rngHashState", "rngTeaState"):
        text = re.sub(r"\bstatic\s+\w+\s+" + name + r"\s*;", "", text)
        text = re.sub(r"\b" + name + r"\b", f"hlsl::currentFiber().{name}", text)

    # DXC evaluates call arguments left to right, C++ leaves the order open ( g++ goes right to left ): the only calls in the shaders whose
    # arguments have side effects are ml.hlsli's uint2( GetUint( ), GetUint( ) ) / uint4( GetUint2( ), GetUint2( ) ) -> braces fix the order
    text = re.sub(r"\b(uint[24])\s*\(\s*(GetUint2?\s*\([^()]*\))\s*,\s*(GetUint2?\s*\([^()]*\))\s*\)", r"\1{ \2, \3 }", text)

    # groupshared T name[ .. ][ .. ]; -> thread_local storage ( one OS thread executes a group ) that the runtime zero-fills per group
    text = re.sub(r"\bgroupshared\s+([A-Za-z_]\w*)\s+(\w+)\s*((?:\[[^\]]*\]\s*)*);",
                  lambda k: f"static thread_local {k.group(1)} {k.group(2)}{k.group(3)}; static hlsl::SmemRegistrar _smem_{k.group(2)}( &g_module, []() -> void* {{ return (void*)&{k.group(2)}; }}, sizeof( {k.group(2)} ) );",
                  text)
    if re.search(r"\bgroupshared\b", text):
        raise SystemExit(f"{identifier}: unparsed groupshared declaration")
    # ( Type )0 zero-initialisation of structs
    text = re.sub(r"=\s*\(\s*([A-Z]\w*)\s*\)\s*0\s*;", r"= \1();", text)

    # test probe ( tests/test_parity_at_baseline_sizes_gpu.py ): count how often the spatial passes take the "tap was mirrored" branch of
    # REBLUR_Common_SpatialFilter.hlsli:198 — the predicate hangs on the last mantissa bit of the tap position ( DESIGN.md "chaotic predicates" ),
    # so the CUDA kernels are compared on its RATE. The value of the expression is passed through unchanged.
    text = re.sub(r"any\s*\(\s*uv\s*!=\s*mirrorUv\s*\)", "hlsl::probeMirror( any( uv != mirrorUv ) )", text)

    # HLSL promotes `uint % float` to a float remainder ( REBLUR_Validation.cs.hlsl:215 ); C++ has no such operator
    text = re.sub(r"\(\s*gFrameIndex\s*>>\s*2\s*\)\s*%\s*gMaxAccumulatedFrameNum", "uint( fmod( float( gFrameIndex >> 2 ), gMaxAccumulatedFrameNum ) )", text)

    use_fibers = bool(re.search(r"\bGroupMemoryBarrier(WithGroupSync)?\b|\bQuadRead\w+\b", text))
    x, y, z = (threads + [1, 1])[:3]
    return f"""#include "hlsl_cpu.h"
namespace hlsl {{ namespace {namespace} {{
static ShaderModule g_module;
{text}
static void hlsl_entry( uint3 gtid, uint3 gid, uint3 dtid, uint gi ) {{ (void)gtid; (void)gid; (void)dtid; (void)gi; hlsl_main( {", ".join(args)} ); }}
static struct Registrar {{ Registrar() {{ registerShader( ShaderEntry{{ "{identifier}", &g_module, hlsl_entry, {x}u, {y}u, {z}u, {"true" if use_fibers else "false"} }} ); }} }} g_registrar;
}} }}
"""


if __name__ == "__main__":
    sys.stdout.write(transform(sys.stdin.read(), sys.argv[1], sys.argv[2]))
