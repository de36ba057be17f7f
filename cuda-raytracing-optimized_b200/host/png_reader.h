// png_reader.h -- PNG file / memory image -> 8-bit RGB rows (top-down, or bottom-up with flipVertically), the decode step of the
// reference's loadTexture (staircase_scene.h:103-118). Returns false for anything that is not a well-formed PNG.
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

namespace crt {
bool decodePng(const uint8_t* bytes, size_t n, bool flipVertically, int& width, int& height, std::vector<uint8_t>& rgb);
bool readPngFile(const char* path, bool flipVertically, int& width, int& height, std::vector<uint8_t>& rgb);
} // namespace crt
