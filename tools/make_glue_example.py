"""examples/p64gpu_glue.c = the C blocks of INTEGRATION.md sections 3, 3b and 3c in one file, so that the binding the
document shows is a file a compiler has seen (tests syntax-check it against the reference's own headers)."""
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
s = open(os.path.join(ROOT, "INTEGRATION.md")).read()


def block(start):
    a = s.index("```c\n" + start)
    return s[a + 5:s.index("```", a + 5)]


out = '''/* Reference-side binding of maikmerten/p64 to libp64b200.so -- the file a maintainer adds to the reference tree
 * (INTEGRATION.md sections 3, 3b, 3c show it piece by piece; generated from them by tools/make_glue_example.py).  It uses
 * the reference's own headers (globals.h) and globals; tests/test_abi_and_host.py syntax-checks it against them when the
 * reference tree is present. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
''' + block('#include "globals.h"') + '''
/* ---- 3b: the device writes the bits too (fixed quantiser) ---- */
static unsigned int pending, pending_len;
extern int FirstFrameBits, NumberOvfl, FrameRate, FrameRateDiv, FrameSkip, QDFact, QOffs;
''' + block('void p64gpu_frame_bits(void)') + '''
/* ---- 3c: rate control without the round trips ---- */
''' + block('void p64gpu_init_rate(void)')
open(os.path.join(ROOT, "examples", "p64gpu_glue.c"), "w").write(out)
