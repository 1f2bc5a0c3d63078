// Placeholders until K2 (body mask) and K7 (label clean-up) land.
#include "common.cuh"
extern "C" size_t eitb_label_cleanup_workspace_bytes(int B, int H, int W) { (void)B; (void)H; (void)W; return 0; }
extern "C" int eitb_label_cleanup(uint8_t*, const uint8_t*, int, int, int, void*, size_t, eitb_stream_t) { return EITB_ERR_UNSUPPORTED; }
