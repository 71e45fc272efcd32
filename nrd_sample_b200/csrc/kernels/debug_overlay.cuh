// Shared pieces of the validation overlays ( CommonSettings::enableValidation, NRDSettings.h:193; REBLUR_Validation.cs.hlsl, RELAX_Validation.cs.hlsl ):
// a 4 x 4 grid of viewports over OUT_VALIDATION, each showing one input or internal quantity with a caption in its top-left corner.
//
// Captions: the reference prints them with MathLib's 5x6 bitmap font ( ml.hlsli "Text" ). The layout here is the same — glyph cells of ( 5 + 1 ) x 6 pixels
// starting 5 pixels into the viewport, digits right to left as Print_ui does — so captions sit where an NRD user expects them; the glyph bitmaps are this
// file's own, so caption pixels are NOT compared with the reference ( the parity test masks the caption rows ), everything else is.
#pragma once
#include "reblur_common.cuh"

namespace nrdk {

// 5x6 glyphs, one 5-bit row per hex digit pair, top row first, bit 4 = left pixel. 0-9, '-', '.', ' ', '&', A-Z
__device__ constexpr unsigned char kGlyphRows[40][6] = {
    {0x0E, 0x11, 0x13, 0x15, 0x19, 0x0E}, {0x04, 0x0C, 0x04, 0x04, 0x04, 0x0E}, {0x0E, 0x11, 0x02, 0x04, 0x08, 0x1F}, {0x1E, 0x01, 0x0E, 0x01, 0x01, 0x1E},
    {0x12, 0x12, 0x1F, 0x02, 0x02, 0x02}, {0x1F, 0x10, 0x1E, 0x01, 0x01, 0x1E}, {0x0E, 0x10, 0x1E, 0x11, 0x11, 0x0E}, {0x1F, 0x01, 0x02, 0x04, 0x04, 0x04},
    {0x0E, 0x11, 0x0E, 0x11, 0x11, 0x0E}, {0x0E, 0x11, 0x11, 0x0F, 0x01, 0x0E}, {0x00, 0x00, 0x1F, 0x00, 0x00, 0x00}, {0x00, 0x00, 0x00, 0x00, 0x0C, 0x0C},
    {0x00, 0x00, 0x00, 0x00, 0x00, 0x00}, {0x0C, 0x12, 0x0C, 0x15, 0x12, 0x0D},
    {0x0E, 0x11, 0x11, 0x1F, 0x11, 0x11}, {0x1E, 0x11, 0x1E, 0x11, 0x11, 0x1E}, {0x0F, 0x10, 0x10, 0x10, 0x10, 0x0F}, {0x1E, 0x11, 0x11, 0x11, 0x11, 0x1E},
    {0x1F, 0x10, 0x1E, 0x10, 0x10, 0x1F}, {0x1F, 0x10, 0x1E, 0x10, 0x10, 0x10}, {0x0F, 0x10, 0x13, 0x11, 0x11, 0x0F}, {0x11, 0x11, 0x1F, 0x11, 0x11, 0x11},
    {0x0E, 0x04, 0x04, 0x04, 0x04, 0x0E}, {0x01, 0x01, 0x01, 0x01, 0x11, 0x0E}, {0x11, 0x12, 0x1C, 0x12, 0x11, 0x11}, {0x10, 0x10, 0x10, 0x10, 0x10, 0x1F},
    {0x11, 0x1B, 0x15, 0x11, 0x11, 0x11}, {0x11, 0x19, 0x15, 0x13, 0x11, 0x11}, {0x0E, 0x11, 0x11, 0x11, 0x11, 0x0E}, {0x1E, 0x11, 0x11, 0x1E, 0x10, 0x10},
    {0x0E, 0x11, 0x11, 0x15, 0x12, 0x0D}, {0x1E, 0x11, 0x11, 0x1E, 0x12, 0x11}, {0x0F, 0x10, 0x0E, 0x01, 0x01, 0x1E}, {0x1F, 0x04, 0x04, 0x04, 0x04, 0x04},
    {0x11, 0x11, 0x11, 0x11, 0x11, 0x0E}, {0x11, 0x11, 0x11, 0x0A, 0x0A, 0x04}, {0x11, 0x11, 0x11, 0x15, 0x1B, 0x11}, {0x11, 0x0A, 0x04, 0x04, 0x0A, 0x11},
    {0x11, 0x0A, 0x04, 0x04, 0x04, 0x04}, {0x1F, 0x02, 0x04, 0x08, 0x10, 0x1F},
};
constexpr int kGlyphW = 5, kGlyphH = 6, kGlyphMinus = 10, kGlyphDot = 11, kGlyphSpace = 12, kGlyphAmpersand = 13, kGlyphA = 14;

// Cursor semantics of ml.hlsli Text ( Init / Print_ch / NextChar / NextDigit / Print_ui ): x, y are this pixel's position relative to the next glyph's origin
struct Caption {
    unsigned x, y;
    bool foreground;
    NRD_DEV Caption(int px, int py, int originX, int originY) : x((unsigned)(px - originX)), y((unsigned)(py - originY)), foreground(false) {}
    NRD_DEV void render(int glyph) {
        if (x < (unsigned)kGlyphW && y < (unsigned)kGlyphH) foreground = ((kGlyphRows[glyph][y] >> (kGlyphW - 1 - x)) & 1u) != 0u;
    }
    NRD_DEV void nextChar() { x -= kGlyphW + 1; }
    NRD_DEV void nextDigit() { x += kGlyphW + 1; }
    NRD_DEV void print(const char* text) {   // upper-case letters, '-', '.', ' ', '&'
        for (; *text; ++text) {
            const char c = *text;
            render(c == '-' ? kGlyphMinus : (c == '.' ? kGlyphDot : (c == ' ' ? kGlyphSpace : (c == '&' ? kGlyphAmpersand : kGlyphA + (c - 'A')))));
            nextChar();
        }
    }
    NRD_DEV void printUint(unsigned v) {     // least significant digit first, walking right ( Print_ui )
        while (v) {
            const unsigned q = v / 10u;
            nextDigit();
            render((int)(v - q * 10u));
            v = q;
        }
    }
};

// Color::ColorizeZucconi ( ml.hlsli:1147-1167 ): spectral ramp used for the "accumulated frames" viewports
NRD_DEV float3 colorizeZucconi(float x) {
    x = saturate(x) * 0.85f;
    const float3 c1 = make_float3(3.54585104f, 2.93225262f, 2.41593945f), x1 = make_float3(0.69549072f, 0.49228336f, 0.27699880f), y1 = make_float3(0.02312639f, 0.15225084f, 0.52607955f);
    const float3 c2 = make_float3(3.90307140f, 3.21182957f, 3.96587128f), x2 = make_float3(0.11748627f, 0.86755042f, 0.66077860f), y2 = make_float3(0.84897130f, 0.88445281f, 0.73949448f);
    const float3 t = c1 * (x - x1), k = c2 * (x - x2);
    const float3 a = make_float3(saturate(1.0f - t.x * t.x - y1.x), saturate(1.0f - t.y * t.y - y1.y), saturate(1.0f - t.z * t.z - y1.z));
    const float3 b = make_float3(saturate(1.0f - k.x * k.x - y2.x), saturate(1.0f - k.y * k.y - y2.y), saturate(1.0f - k.z * k.z - y2.z));
    return make_float3(saturate(a.x + b.x), saturate(a.y + b.y), saturate(a.z + b.z));
}

// Where the pixel sits in the 4 x 4 grid ( REBLUR_Validation.cs.hlsl:45-51 )
struct OverlayCell {
    float2 uv;        // position inside the viewport, 0..1
    int idx, idy;     // viewport column / row
    int index;        // row * 4 + column
    int captionX, captionY;
};
NRD_DEV OverlayCell overlayCell(int px, int py, float2 resourceSize) {
    // every viewport uv lands EXACTLY on a texel edge of the full-size inputs ( ( px + 0.5 ) * 4 is an integer ), so which texel the point sampling picks is decided
    // by the rounding of this division: IEEE, not the approximate one -use_fast_math substitutes
    const float2 pixelUv = make_float2(__fdiv_rn((float)px + 0.5f, resourceSize.x), __fdiv_rn((float)py + 0.5f, resourceSize.y));
    const float2 scaled = pixelUv / 0.25f;
    OverlayCell c;
    c.uv = frac2(scaled);
    const float2 id = floor2(scaled);
    c.idx = (int)id.x;
    c.idy = (int)id.y;
    c.index = (int)(id.y / 0.25f + id.x);
    c.captionX = (int)(unsigned)(id.x * resourceSize.x * 0.25f + 5.0f);   // uint2( viewportId * gResourceSize * VIEWPORT_SIZE + OFFSET )
    c.captionY = (int)(unsigned)(id.y * resourceSize.y * 0.25f + 5.0f);
    return c;
}
// the final touch every viewport gets: caption pixels invert the picture underneath ( :329-333 )
NRD_DEV float3 applyCaption(float3 rgb, bool foreground) {
    if (!foreground) return rgb;
    const float lum = rgb.x * 0.2126f + rgb.y * 0.7152f + rgb.z * 0.0722f;   // Color::Luminance ( ml.hlsli:712-717 )
    const float t = saturate(fabsf(lum - 0.5f) / 0.25f);
    return make_float3((1.0f - rgb.x) * t, (1.0f - rgb.y) * t, (1.0f - rgb.z) * t);
}

}  // namespace nrdk
