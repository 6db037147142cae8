/* TEST INFRASTRUCTURE ONLY (oracle/): force-included when compiling the reference's
 * libbch/bch.c so that its hard-wired LOGGING printf calls (bch.c:21, ~20 per decode, one
 * unconditional at :277) do not dominate CPU-baseline timings.  The reference source is
 * not modified. */
#include <stdio.h>
#define printf(...) ((void)0)
