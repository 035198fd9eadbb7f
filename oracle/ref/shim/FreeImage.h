// Stand-in for <FreeImage.h> — TEST INFRASTRUCTURE ONLY.  The reference's util/bitmap.h (included
// by base/reconstruction.cc) names these; nothing on the tested path touches an image file.
#pragma once
struct FIBITMAP;
enum FREE_IMAGE_FORMAT { FIF_UNKNOWN = -1 };
enum FREE_IMAGE_FILTER { FILTER_BOX = 0, FILTER_BICUBIC = 1, FILTER_BILINEAR = 2 };
enum FREE_IMAGE_MDMODEL { FIMD_NODATA = -1, FIMD_EXIF_MAIN = 1, FIMD_EXIF_EXIF = 2, FIMD_EXIF_GPS = 3 };
extern "C" {
void FreeImage_Unload(FIBITMAP*);
unsigned FreeImage_GetBPP(FIBITMAP*);
unsigned FreeImage_GetPitch(FIBITMAP*);
unsigned FreeImage_GetWidth(FIBITMAP*);
unsigned FreeImage_GetHeight(FIBITMAP*);
}
