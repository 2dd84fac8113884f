/*
 * pll_maps.c - character -> state-mask tables (reference src/maps.c:26-110).
 *
 * State i of an alphabet is bit i of the mask; ambiguity codes are unions.  Written as
 * designated initialisers from the IUPAC definitions rather than as 256-entry dumps.
 *   nucleotides : A=1 C=2 G=4 T/U=8, IUPAC unions, gap/unknown (- ? N O X) = 15
 *   amino acids : A R N D C Q E G H I L K M F P S T W Y V = bits 0..19,
 *                 B = N|D, Z = Q|E, gap/unknown (- ? * X) = all 20 bits
 *   binary      : 0 -> 1, 1 -> 2, - ? -> 3
 */
#include "pll.h"

#define BOTH(ch, v) [ch] = (v), [(ch) + 32] = (v) /* upper and lower case */

PLL_EXPORT const unsigned int pll_map_bin[256] = {
  ['0'] = 1, ['1'] = 2, ['-'] = 3, ['?'] = 3,
};

#define NT_A 1u
#define NT_C 2u
#define NT_G 4u
#define NT_T 8u

PLL_EXPORT const unsigned int pll_map_nt[256] = {
  ['-'] = 15, ['?'] = 15,
  BOTH('A', NT_A), BOTH('C', NT_C), BOTH('G', NT_G), BOTH('T', NT_T), BOTH('U', NT_T),
  BOTH('M', NT_A | NT_C), BOTH('R', NT_A | NT_G), BOTH('W', NT_A | NT_T),
  BOTH('S', NT_C | NT_G), BOTH('Y', NT_C | NT_T), BOTH('K', NT_G | NT_T),
  BOTH('V', NT_A | NT_C | NT_G), BOTH('H', NT_A | NT_C | NT_T),
  BOTH('D', NT_A | NT_G | NT_T), BOTH('B', NT_C | NT_G | NT_T),
  BOTH('N', 15), BOTH('O', 15), BOTH('X', 15),
};

#define AA(i) (1u << (i))
#define AA_ANY 0xfffffu

PLL_EXPORT const unsigned int pll_map_aa[256] = {
  ['*'] = AA_ANY, ['-'] = AA_ANY, ['?'] = AA_ANY,
  BOTH('A', AA(0)),  BOTH('R', AA(1)),  BOTH('N', AA(2)),  BOTH('D', AA(3)),  BOTH('C', AA(4)),
  BOTH('Q', AA(5)),  BOTH('E', AA(6)),  BOTH('G', AA(7)),  BOTH('H', AA(8)),  BOTH('I', AA(9)),
  BOTH('L', AA(10)), BOTH('K', AA(11)), BOTH('M', AA(12)), BOTH('F', AA(13)), BOTH('P', AA(14)),
  BOTH('S', AA(15)), BOTH('T', AA(16)), BOTH('W', AA(17)), BOTH('Y', AA(18)), BOTH('V', AA(19)),
  BOTH('B', AA(2) | AA(3)), BOTH('Z', AA(5) | AA(6)), BOTH('X', AA_ANY),
};
