// oracle/verifiers_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// The two remaining independent PNG decoders of the reference's round-trip check (src/fpng_test.cpp:1403-1445 wuffs,
// 1571-1606 pvpng; lodepng and stb_image are in ref_shim.cpp) behind extern "C" wrappers: the UNMODIFIED third-party sources
// are compiled from /root/reference/src where they lie (wuffs-v0.3.c, pvpngreader.cpp + basisu_miniz.h) by oracle/Makefile
// into oracle/_ref/libpng_verifiers.so; nothing is copied into this repository.
//
// Unlike the harness (which switches wuffs's checksum verification off, fpng_test.cpp:690), wuffs runs here with its
// defaults: it verifies the IDAT CRC-32 and the zlib Adler-32, so a file it accepts has both checksums right.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define WUFFS_IMPLEMENTATION
#define WUFFS_CONFIG__STATIC_FUNCTIONS
#include "wuffs-v0.3.c"
#include "pvpngreader.h"
#include "basisu_miniz.h"          // the miniz implementation pvpngreader.cpp links against (the harness instantiates it the same way, fpng_test.cpp:30)

#define VER_API extern "C" __attribute__((visibility("default")))

// Decodes to 8-bit RGBA (non-premultiplied) into out[0 .. 4*w*h).  Returns 0 on success, 1 = not decodable / checksum mismatch,
// 2 = out_cap too small (w and h are set), 3 = allocation failure.
VER_API int ver_wuffs_decode_rgba(const void* file, size_t size, int ignore_checksums, void* out, size_t out_cap, uint32_t* w, uint32_t* h)
{
    *w = *h = 0;
    wuffs_png__decoder* dec = wuffs_png__decoder__alloc();          // zero-initialised + initialised decoder object
    if (!dec) return 3;
    int rc = 1;
    uint8_t* work = NULL;
    do {
        if (ignore_checksums) wuffs_png__decoder__set_quirk_enabled(dec, WUFFS_BASE__QUIRK_IGNORE_CHECKSUM, true);
        wuffs_base__io_buffer src = wuffs_base__ptr_u8__reader((uint8_t*)file, size, /*closed=*/true);
        wuffs_base__image_config cfg;
        if (wuffs_png__decoder__decode_image_config(dec, &cfg, &src).repr) break;
        const uint32_t ww = wuffs_base__pixel_config__width(&cfg.pixcfg), hh = wuffs_base__pixel_config__height(&cfg.pixcfg);
        *w = ww; *h = hh;
        const uint64_t need = (uint64_t)ww * hh * 4u;
        if (need > out_cap) { rc = 2; break; }
        wuffs_base__pixel_config__set(&cfg.pixcfg, WUFFS_BASE__PIXEL_FORMAT__RGBA_NONPREMUL, WUFFS_BASE__PIXEL_SUBSAMPLING__NONE, ww, hh);
        const uint64_t work_len = wuffs_png__decoder__workbuf_len(dec).max_incl;
        work = (uint8_t*)malloc(work_len ? (size_t)work_len : 1);
        if (!work) { rc = 3; break; }
        wuffs_base__pixel_buffer pb;
        if (wuffs_base__pixel_buffer__set_from_slice(&pb, &cfg.pixcfg, wuffs_base__make_slice_u8((uint8_t*)out, (size_t)need)).repr) break;
        if (wuffs_png__decoder__decode_frame(dec, &pb, &src, WUFFS_BASE__PIXEL_BLEND__SRC,
                                             wuffs_base__make_slice_u8(work, (size_t)work_len), NULL).repr) break;
        rc = 0;
    } while (0);
    free(work);
    free(dec);
    return rc;
}

// pvpng (the PNG reader that ships with Basis Universal; inflates with miniz).  desired_chans 3 or 4.  Returns 0 on success,
// 1 = not decodable, 2 = out_cap too small; *chans_in_file as pvpng reports it.
VER_API int ver_pvpng_decode(const void* file, size_t size, uint32_t desired_chans, void* out, size_t out_cap,
                             uint32_t* w, uint32_t* h, uint32_t* chans_in_file)
{
    *w = *h = *chans_in_file = 0;
    uint32_t ww = 0, hh = 0, cc = 0;
    void* p = pv_png::load_png(file, size, desired_chans, ww, hh, cc);
    if (!p) return 1;
    *w = ww; *h = hh; *chans_in_file = cc;
    const uint64_t need = (uint64_t)ww * hh * desired_chans;
    const int rc = need <= out_cap ? 0 : 2;
    if (rc == 0) memcpy(out, p, (size_t)need);
    free(p);
    return rc;
}
