/* pregraph_shim.c -- the ONE object a maintainer compiles into SOAPdenovo-{63,127}mer in place of pregraph.c (and of the five
 * files only pregraph.c calls into: prlHashReads.c cutTipPreGraph.c node2edge.c prlRead2path.c output_pregraph.c).
 * It provides the symbol the reference's main.c binds (main.c:29 `extern int call_pregraph(int, char **)`, called at main.c:74 for
 * the `pregraph` sub-command and main.c:341 inside `all`) and fixes the 63-mer / 127-mer flavour at BUILD time, exactly as the
 * reference does with -DMER63 / -DMER127 (standardPregraph/Makefile:51-66).  scripts/link_dropin.sh builds both flavours this way.
 */
#include "../../include/pregraph_b200.h"
#ifndef PGB_FLAVOUR127
#error "compile with -DPGB_FLAVOUR127=0 (SOAPdenovo-63mer) or -DPGB_FLAVOUR127=1 (SOAPdenovo-127mer)"
#endif
int call_pregraph(int argc, char **argv) { return pgb200_pregraph_main(argc, argv, PGB_FLAVOUR127); }
