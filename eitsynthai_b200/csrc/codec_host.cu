// Host-side decoders for the two lossless compressed transfer syntaxes CT scanners and PACS export most often
// (row (f-1) of SURVEY section 8, "ingest"): the reference leaves them to pydicom + pylibjpeg
// (kt_service/ai_tools/utils.py:52-60 dcmread, requirements.txt:9-13); this library decodes them itself so that a
// compressed series reaches the pinned upload buffer without those packages.
//   * RLE Lossless            1.2.840.10008.1.2.5   (DICOM PS3.5 Annex G: byte planes, PackBits runs)
//   * JPEG Lossless, SV1 / any selection value  1.2.840.10008.1.2.4.70 / .57  (ITU-T T.81 process 14: Huffman-coded
//     prediction differences, one component, 2..16 bits, restart intervals)
// Plain C-ABI, host pointers in and out, no CUDA: compiled into libeitb200 with the kernels.
#include <stdint.h>
#include <string.h>
#include "../../include/eitb200.h"

namespace {

// ------------------------------------------------------------------------------------------------ PackBits
// returns bytes written, or -1 on a malformed run
long packbits(const uint8_t* src, size_t n, uint8_t* dst, size_t cap) {
    size_t i = 0, o = 0;
    while (i < n && o < cap) {
        const int c = (int8_t)src[i++];
        if (c >= 0) {
            size_t cnt = (size_t)c + 1;
            if (i + cnt > n) return -1;
            if (cnt > cap - o) cnt = cap - o;
            memcpy(dst + o, src + i, cnt);
            i += (size_t)c + 1; o += cnt;
        } else if (c != -128) {
            size_t cnt = (size_t)(1 - c);
            if (i >= n) return -1;
            if (cnt > cap - o) cnt = cap - o;
            memset(dst + o, src[i], cnt);
            ++i; o += cnt;
        }
    }
    return (long)o;
}

// ------------------------------------------------------------------------------------------------ JPEG lossless
struct Huff {
    int present;
    int mincode[17], maxcode[18], valptr[17];
    uint8_t vals[256];
};

void build_huff(Huff* h, const uint8_t* bits /*16*/, const uint8_t* vals, int nvals) {
    int code = 0, k = 0;
    for (int l = 1; l <= 16; ++l) {
        h->valptr[l] = k;
        h->mincode[l] = code;
        code += bits[l - 1];
        k += bits[l - 1];
        h->maxcode[l] = bits[l - 1] ? code - 1 : -1;
        code <<= 1;
    }
    h->maxcode[17] = 0x7fffffff;
    memcpy(h->vals, vals, (size_t)nvals);
    h->present = 1;
}

struct Bits {
    const uint8_t* p;
    const uint8_t* end;
    uint32_t acc;
    int n;
    int hit_marker;
};

inline void refill(Bits* b) {
    while (b->n <= 24) {
        uint32_t v = 0;
        if (!b->hit_marker && b->p < b->end) {
            v = *b->p;
            if (v == 0xFF) {
                const int nx = b->p + 1 < b->end ? b->p[1] : 0xD9;
                if (nx == 0) b->p += 2;                       // stuffed zero
                else { b->hit_marker = 1; v = 0; }            // a marker: feed zeros until the caller deals with it
            } else {
                ++b->p;
            }
        }
        b->acc |= v << (24 - b->n);
        b->n += 8;
    }
}
inline int getbit(Bits* b) {
    if (b->n == 0) refill(b);
    const int r = (int)(b->acc >> 31);
    b->acc <<= 1; --b->n;
    return r;
}
inline int getbits(Bits* b, int k) {
    if (k == 0) return 0;
    if (b->n < k) refill(b);
    const int r = (int)(b->acc >> (32 - k));
    b->acc <<= k; b->n -= k;
    return r;
}
inline int decode_symbol(Bits* b, const Huff* h) {
    int code = 0;
    for (int l = 1; l <= 16; ++l) {
        code = (code << 1) | getbit(b);
        if (h->maxcode[l] >= 0 && code <= h->maxcode[l] && code >= h->mincode[l]) return h->vals[h->valptr[l] + code - h->mincode[l]];
    }
    return -1;
}

}  // namespace

/* One RLE-compressed frame (the fragment that follows the basic offset table) -> rows*cols samples of
 * bytes_per_sample bytes, little endian. */
extern "C" int eitb_rle_decode_frame(const uint8_t* frag, size_t frag_len, int rows, int cols, int bytes_per_sample, uint8_t* out) {
    if (!frag || !out || rows <= 0 || cols <= 0 || bytes_per_sample < 1 || bytes_per_sample > 4 || frag_len < 64) return EITB_ERR_BAD_ARG;
    uint32_t hdr[16];
    memcpy(hdr, frag, 64);
    const int nseg = (int)hdr[0];
    if (nseg != bytes_per_sample) return EITB_ERR_UNSUPPORTED;   // one sample per pixel: one segment per byte plane
    const size_t npx = (size_t)rows * cols;
    uint8_t* plane = new uint8_t[npx];
    int rc = EITB_OK;
    for (int s = 0; s < nseg && rc == EITB_OK; ++s) {
        const size_t a = hdr[1 + s], z = s + 1 < nseg ? hdr[2 + s] : frag_len;
        if (a < 64 || a > z || z > frag_len) { rc = EITB_ERR_BAD_ARG; break; }
        const long got = packbits(frag + a, z - a, plane, npx);
        if (got != (long)npx) { rc = EITB_ERR_BAD_ARG; break; }
        const int byte_pos = nseg - 1 - s;                        // segment 0 is the most significant byte
        for (size_t i = 0; i < npx; ++i) out[i * nseg + byte_pos] = plane[i];
    }
    delete[] plane;
    return rc;
}

/* A JPEG lossless (SOF3) bit stream -> out[rows*cols] uint16 (sample values, point transform undone).
 * rows / cols / precision are reported; out may be NULL to query them. */
extern "C" int eitb_jpeg_lossless_decode(const uint8_t* data, size_t len, int* rows_out, int* cols_out, int* precision_out, uint16_t* out,
                                         size_t out_capacity) {
    if (!data || len < 4 || data[0] != 0xFF || data[1] != 0xD8) return EITB_ERR_BAD_ARG;
    Huff tables[4];
    memset(tables, 0, sizeof(tables));
    int P = 0, Y = 0, X = 0, restart = 0;
    size_t pos = 2;
    while (pos + 4 <= len) {
        if (data[pos] != 0xFF) return EITB_ERR_BAD_ARG;
        const int m = data[pos + 1];
        if (m == 0xFF) { ++pos; continue; }                       // fill byte
        const size_t seg = ((size_t)data[pos + 2] << 8) | data[pos + 3];
        const uint8_t* q = data + pos + 4;
        if (pos + 2 + seg > len) return EITB_ERR_BAD_ARG;
        if (m == 0xC4) {                                           // DHT
            size_t o = 0;
            while (o + 17 <= seg - 2) {
                const int th = q[o] & 15;
                int nv = 0;
                for (int i = 0; i < 16; ++i) nv += q[o + 1 + i];
                if (th > 3 || nv > 256 || o + 17 + (size_t)nv > seg - 2) return EITB_ERR_BAD_ARG;
                build_huff(&tables[th], q + o + 1, q + o + 17, nv);
                o += 17 + (size_t)nv;
            }
        } else if (m == 0xC3) {                                    // SOF3: lossless, Huffman
            P = q[0]; Y = (q[1] << 8) | q[2]; X = (q[3] << 8) | q[4];
            if (q[5] != 1) return EITB_ERR_UNSUPPORTED;            // one component (monochrome CT)
            if (P < 2 || P > 16 || Y <= 0 || X <= 0) return EITB_ERR_BAD_ARG;
        } else if (m >= 0xC0 && m <= 0xCF && m != 0xC8 && m != 0xCC) {
            return EITB_ERR_UNSUPPORTED;                           // any other frame type (baseline, JPEG-LS is F7, ...)
        } else if (m == 0xDD) {                                    // DRI
            restart = (q[0] << 8) | q[1];
        } else if (m == 0xDA) {                                    // SOS
            if (!P) return EITB_ERR_BAD_ARG;
            if (rows_out) *rows_out = Y;
            if (cols_out) *cols_out = X;
            if (precision_out) *precision_out = P;
            if (!out) return EITB_OK;
            if (out_capacity < (size_t)Y * X) return EITB_ERR_WORKSPACE;
            if (q[0] != 1) return EITB_ERR_UNSUPPORTED;
            const int td = q[2] >> 4, sel = q[3], pt = q[5] & 15;
            if (sel < 1 || sel > 7 || !tables[td].present) return EITB_ERR_UNSUPPORTED;
            const Huff* h = &tables[td];
            Bits br = {data + pos + 2 + seg, data + len, 0u, 0, 0};
            const int def = 1 << (P - pt - 1);
            int since_restart = 0, rst_row = 0;                    // rst_row: the row that began (or continues) a restart interval
            bool fresh = true;                                     // next sample starts an interval
            for (int y = 0; y < Y; ++y) {
                uint16_t* row = out + (size_t)y * X;
                const uint16_t* up = y > 0 ? row - X : nullptr;
                for (int x = 0; x < X; ++x) {
                    if (restart && since_restart == restart) {     // RSTn expected here
                        br.n = 0; br.acc = 0;                      // drop the padding bits
                        if (br.hit_marker || (br.p + 1 < br.end && br.p[0] == 0xFF && br.p[1] >= 0xD0 && br.p[1] <= 0xD7)) {
                            br.p += 2; br.hit_marker = 0;
                        } else {
                            return EITB_ERR_BAD_ARG;
                        }
                        since_restart = 0; fresh = true;
                    }
                    const int s = decode_symbol(&br, h);
                    if (s < 0 || s > 16) return EITB_ERR_BAD_ARG;
                    int diff;
                    if (s == 0) diff = 0;
                    else if (s == 16) diff = 32768;
                    else { const int v = getbits(&br, s); diff = v < (1 << (s - 1)) ? v - (1 << s) + 1 : v; }
                    int pred;
                    if (fresh) { pred = def; rst_row = y; fresh = false; }
                    else if (y == rst_row || !up) pred = x > 0 ? (row[x - 1] >> pt) : (up ? (up[0] >> pt) : def);   // first line of an interval: Ra
                    else if (x == 0) pred = up[0] >> pt;                                                           // first column: Rb
                    else {
                        const int ra = row[x - 1] >> pt, rb = up[x] >> pt, rc = up[x - 1] >> pt;
                        switch (sel) {
                            case 1: pred = ra; break;
                            case 2: pred = rb; break;
                            case 3: pred = rc; break;
                            case 4: pred = ra + rb - rc; break;
                            case 5: pred = ra + ((rb - rc) >> 1); break;
                            case 6: pred = rb + ((ra - rc) >> 1); break;
                            default: pred = (ra + rb) >> 1; break;
                        }
                    }
                    row[x] = (uint16_t)(((pred + diff) & 0xffff) << pt);
                    ++since_restart;
                }
            }
            return EITB_OK;
        }
        pos += 2 + seg;
    }
    return EITB_ERR_BAD_ARG;
}
