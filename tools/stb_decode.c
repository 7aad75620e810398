/* stb_decode — decode one image to raw RGBA8 exactly as the reference's importer sees it.
 *
 * The reference loads glTF images through tinygltf, which calls stb_image with 4 components
 * forced (Foreground/SceneGraph/tiny_gltf.h:1676-1766, asserted in
 * Foreground/SceneGraph/glTFSceneImporter.cpp:116-117).  JPEG decoders differ in their IDCT and
 * chroma upsampling, so to get the reference's bytes we compile against the reference's own
 * stb_image.h *where it lies* (never copied into this repo):
 *
 *   gcc -O2 -I/root/reference/Foreground/SceneGraph tools/stb_decode.c -lm -o oracle/_ref/stb_decode
 *
 * usage: stb_decode in.jpg out.rgba   → writes "W H\n" to stdout and W*H*4 bytes to out.rgba
 */
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"
#include <stdio.h>

int main(int argc, char** argv)
{
    if (argc != 3) { fprintf(stderr, "usage: %s in out.rgba\n", argv[0]); return 2; }
    int w, h, comp;
    unsigned char* px = stbi_load(argv[1], &w, &h, &comp, 4);
    if (!px) { fprintf(stderr, "stb_decode: cannot load %s: %s\n", argv[1], stbi_failure_reason()); return 1; }
    FILE* f = fopen(argv[2], "wb");
    if (!f) { perror("fopen"); return 1; }
    fwrite(px, 1, (size_t)w * h * 4, f);
    fclose(f);
    printf("%d %d\n", w, h);
    stbi_image_free(px);
    return 0;
}
