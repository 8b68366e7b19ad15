// Host check of the netpbm reader of popsift-demo: prints "w h" and the gray bytes (hex) of one file.
#include "pgmread.h"
#include <cstdio>
int main(int argc, char** argv)
{
    if (argc < 2) return 2;
    int w = 0, h = 0;
    unsigned char* d = readPGMfile(argv[1], w, h);
    if (!d) { std::printf("FAIL\n"); return 1; }
    std::printf("%d %d\n", w, h);
    for (int i = 0; i < w * h; ++i) std::printf("%02x", d[i]);
    std::printf("\n");
    delete[] d;
    return 0;
}
