// tiff_stub.cpp — the raw-container implementation behind oracle/deblock_ref/tiffio.h (test infrastructure only).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "tiffio.h"

struct brief_tiff_stub {
  FILE* f;
  bool writing;
  uint32_t depth, height, width, dir;
  std::vector<uint16_t> data;  // whole volume (read: loaded at open; write: appended page by page)
};

extern "C" {

TIFF* TIFFOpen(const char* name, const char* mode) {
  TIFF* t = new TIFF();
  t->writing = mode[0] == 'w';
  t->dir = 0;
  t->depth = t->height = t->width = 0;
  t->f = fopen(name, t->writing ? "wb" : "rb");
  if (!t->f) { delete t; return nullptr; }
  if (!t->writing) {
    char magic[8];
    uint32_t dims[3];
    if (fread(magic, 1, 8, t->f) != 8 || memcmp(magic, "BRIEFRAW", 8) != 0 || fread(dims, 4, 3, t->f) != 3) {
      fclose(t->f); delete t; return nullptr;
    }
    t->depth = dims[0]; t->height = dims[1]; t->width = dims[2];
    t->data.resize((size_t)dims[0] * dims[1] * dims[2]);
    if (fread(t->data.data(), 2, t->data.size(), t->f) != t->data.size()) { fclose(t->f); delete t; return nullptr; }
  }
  return t;
}

void TIFFClose(TIFF* t) {
  if (t->writing) {
    const uint32_t dims[3] = {t->dir, t->height, t->width};
    fwrite("BRIEFRAW", 1, 8, t->f);
    fwrite(dims, 4, 3, t->f);
    fwrite(t->data.data(), 2, t->data.size(), t->f);
  }
  fclose(t->f);
  delete t;
}

int TIFFGetField(TIFF* t, ttag_t tag, ...) {
  va_list ap;
  va_start(ap, tag);
  if (tag == TIFFTAG_BITSPERSAMPLE) *va_arg(ap, uint16_t*) = 16;
  else if (tag == TIFFTAG_IMAGELENGTH) *va_arg(ap, uint32_t*) = t->height;
  else if (tag == TIFFTAG_IMAGEWIDTH) *va_arg(ap, uint32_t*) = t->width;
  va_end(ap);
  return 1;
}

int TIFFSetField(TIFF* t, ttag_t tag, ...) {
  va_list ap;
  va_start(ap, tag);
  if (tag == TIFFTAG_IMAGEWIDTH) t->width = va_arg(ap, uint32_t);
  else if (tag == TIFFTAG_IMAGELENGTH) t->height = va_arg(ap, uint32_t);
  va_end(ap);
  return 1;
}

uint16_t TIFFNumberOfDirectories(TIFF* t) { return (uint16_t)t->depth; }

int TIFFReadScanline(TIFF* t, void* buf, uint32_t row, uint16_t) {
  memcpy(buf, t->data.data() + ((size_t)t->dir * t->height + row) * t->width, (size_t)t->width * 2);
  return 1;
}
int TIFFReadDirectory(TIFF* t) { return ++t->dir < t->depth; }

int TIFFWriteScanline(TIFF* t, void* buf, uint32_t row, uint16_t) {
  const size_t need = ((size_t)t->dir * t->height + row + 1) * t->width;
  if (t->data.size() < need) t->data.resize(need);
  memcpy(t->data.data() + ((size_t)t->dir * t->height + row) * t->width, buf, (size_t)t->width * 2);
  return 1;
}
int TIFFWriteDirectory(TIFF* t) { ++t->dir; return 1; }

}  // extern "C"
