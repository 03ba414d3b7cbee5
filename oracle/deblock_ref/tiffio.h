/* tiffio.h — a STAND-IN for libtiff's header, just large enough to compile the reference's deblock.cpp
 * (/root/reference/deblock.cpp, its only native component) in a container that has no libtiff.
 * TEST INFRASTRUCTURE ONLY (oracle/): the "TIFF" files this stub reads and writes are a trivial raw container
 *     magic "BRIEFRAW" | uint32 depth, height, width | depth*height*width little-endian uint16 samples
 * produced / consumed by oracle/deblock_oracle.py.  The reference's filter arithmetic, seam discovery and
 * traversal order are compiled from the reference's own source, untouched. */
#ifndef BRIEF_ORACLE_TIFFIO_STUB_H
#define BRIEF_ORACLE_TIFFIO_STUB_H
#include <stdint.h>

typedef struct brief_tiff_stub TIFF;
typedef uint32_t ttag_t;

#define TIFFTAG_SUBFILETYPE 254
#define TIFFTAG_IMAGEWIDTH 256
#define TIFFTAG_IMAGELENGTH 257
#define TIFFTAG_BITSPERSAMPLE 258
#define TIFFTAG_COMPRESSION 259
#define TIFFTAG_SAMPLESPERPIXEL 277
#define TIFFTAG_ROWSPERSTRIP 278
#define TIFFTAG_PAGENUMBER 297
#define FILETYPE_PAGE 2
#define COMPRESSION_NONE 1

#ifdef __cplusplus
extern "C" {
#endif
TIFF* TIFFOpen(const char* name, const char* mode);
void TIFFClose(TIFF* t);
int TIFFGetField(TIFF* t, ttag_t tag, ...);
int TIFFSetField(TIFF* t, ttag_t tag, ...);
uint16_t TIFFNumberOfDirectories(TIFF* t);
int TIFFReadScanline(TIFF* t, void* buf, uint32_t row, uint16_t sample = 0);
int TIFFReadDirectory(TIFF* t);
int TIFFWriteScanline(TIFF* t, void* buf, uint32_t row, uint16_t sample = 0);
int TIFFWriteDirectory(TIFF* t);
#ifdef __cplusplus
}
#endif
#endif
