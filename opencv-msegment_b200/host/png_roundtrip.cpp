// png_roundtrip.cpp -- test helper: decodes an image with the CLI's readers (PNG: png_io.hpp, baseline JPEG: jpeg_io.hpp) and
// re-encodes it as an RGB PNG.   usage: png_roundtrip in.(png|jpg) out.png      (exit 2 = cannot read, 3 = cannot write)
#include "jpeg_io.hpp"
#include "png_io.hpp"

int main(int argc, char** argv)
{
    if (argc != 3) return 1;
    std::vector<uint8_t> bgr;
    int w = 0, h = 0;
    if (!msegment::png::read_bgr(argv[1], bgr, w, h) && !msegment::jpeg::read_bgr(argv[1], bgr, w, h)) return 2;
    for (size_t i = 0; i < (size_t)w * h; i++) std::swap(bgr[3 * i], bgr[3 * i + 2]);
    return msegment::png::write(argv[2], bgr.data(), (size_t)w * 3, w, h, 3) ? 0 : 3;
}
