// H.261 code tables in the form the device-side entropy coder reads (vlc_kernels.cuh): one 32-bit entry per
// symbol, (length << 16) | code bits, 0 = no code.  Filled on the host from the same bit strings the host coder
// uses (vlc_tables.h, bits.cpp), so the two cannot drift apart.
#pragma once
#include <cstdint>

namespace p64b {

struct DevVlcTables {
  uint32_t tcoef[32 * 16];   // [run][|level|] without the sign bit; 0 -> escape (codec.c:113-115)
  uint32_t mtype[10];        // Table 2/H.261, numbered as p64.c:217-222
  uint32_t mvd[32];          // Table 3/H.261, indexed by (difference & 31)
  uint32_t cbp[64];          // Table 4/H.261
  uint32_t pad[2];
};
constexpr int DEV_VLC_WORDS = sizeof(DevVlcTables) / 4;

void fill_dev_vlc_tables(DevVlcTables* d);

}  // namespace p64b
