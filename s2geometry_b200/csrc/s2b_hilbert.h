// s2b_hilbert.h — the two 1024-entry Hilbert lookup tables of S2CellId (s2cell_id.cc:75-115), generated at
// compile time from the curve definition (s2coords_internal.h:44-82): pos[iiiijjjjoo] = ppppppppoo and the
// inverse ij[ppppppppoo] = iiiijjjjoo. Built iteratively (one level = 2 bits) rather than by the reference's recursion.
#pragma once
#include <stdint.h>

namespace s2b {

struct HilbertTables { uint16_t pos[1024]; uint16_t ij[1024]; };

constexpr HilbertTables MakeHilbertTables() {
  // ij -> pos for each orientation (bit0 = swap axes, bit1 = invert bits); child orientation modifier per pos.
  const int ij_to_pos[4][4] = {{0, 1, 3, 2}, {0, 3, 1, 2}, {2, 3, 1, 0}, {2, 1, 3, 0}};
  const int pos_to_orient[4] = {1, 0, 0, 3};
  HilbertTables t{};
  for (int o = 0; o < 4; ++o) {
    for (int i = 0; i < 16; ++i) {
      for (int j = 0; j < 16; ++j) {
        int pos = 0, orient = o;
        for (int b = 3; b >= 0; --b) {
          int ij = (((i >> b) & 1) << 1) | ((j >> b) & 1);
          int p = ij_to_pos[orient][ij];
          pos = (pos << 2) | p;
          orient ^= pos_to_orient[p];
        }
        t.pos[(i << 6) | (j << 2) | o] = (uint16_t)((pos << 2) | orient);
        t.ij[(pos << 2) | o] = (uint16_t)((i << 6) | (j << 2) | orient);
      }
    }
  }
  return t;
}

inline constexpr HilbertTables kHilbert = MakeHilbertTables();

// 6-bit position table (see from_face_ij_wide): out[(i6 << 8) | (j6 << 2) | orientation] = (pos12 << 2) | orientation'.
inline void MakeWideHilbert(uint16_t* out /* 16384 entries */) {
  const int ij_to_pos[4][4] = {{0, 1, 3, 2}, {0, 3, 1, 2}, {2, 3, 1, 0}, {2, 1, 3, 0}};
  const int pos_to_orient[4] = {1, 0, 0, 3};
  for (int o = 0; o < 4; ++o)
    for (int i = 0; i < 64; ++i)
      for (int j = 0; j < 64; ++j) {
        int pos = 0, orient = o;
        for (int b = 5; b >= 0; --b) {
          int ij = (((i >> b) & 1) << 1) | ((j >> b) & 1);
          int p = ij_to_pos[orient][ij];
          pos = (pos << 2) | p;
          orient ^= pos_to_orient[p];
        }
        out[(i << 8) | (j << 2) | o] = (uint16_t)((pos << 2) | orient);
      }
}

}  // namespace s2b
