// ref_stb.cpp — TEST INFRASTRUCTURE. The stb_image the reference vendors and decodes every texture with (Lumen/vendor/stb/stb_image.h;
// LumenPTModelConverter.cpp:121 stbi_load_from_memory(..., 4)), compiled in place by oracle/Makefile into oracle/_ref/ref_stb.
//   ref_stb <image file> <out.rgba>   writes the RGBA8 pixels, prints "width height" (exit 1 when stb cannot decode the file)
#define STB_IMAGE_IMPLEMENTATION
#include <stb/stb_image.h>
#include <cstdio>
int main(int argc, char** argv) {
    if (argc < 3) return 2;
    int w = 0, h = 0, c = 0;
    unsigned char* px = stbi_load(argv[1], &w, &h, &c, 4);
    if (!px) return 1;
    FILE* f = std::fopen(argv[2], "wb"); if (!f) return 2;
    std::fwrite(px, 4, (size_t)w * h, f); std::fclose(f);
    std::printf("%d %d\n", w, h);
    stbi_image_free(px);
    return 0;
}
