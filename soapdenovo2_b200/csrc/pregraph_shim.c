/* pregraph_shim.c -- the ONE object a maintainer compiles into SOAPdenovo-{63,127}mer in place of pregraph.c (and of the five
 * files only pregraph.c calls into: prlHashReads.c cutTipPreGraph.c node2edge.c prlRead2path.c output_pregraph.c).
 * It provides the symbol the reference's main.c binds (main.c:29 `extern int call_pregraph(int, char **)`, called at main.c:74 for
 * the `pregraph` sub-command and main.c:341 inside `all`) and fixes the 63-mer / 127-mer flavour at BUILD time, exactly as the
 * reference does with -DMER63 / -DMER127 (standardPregraph/Makefile:51-66).  scripts/link_dropin.sh builds both flavours this way.
 *
 * It also keeps the process-level contract of the reference's call_pregraph (pregraph.c:62-139) towards the other stages of `all`:
 *   in :  `all` does not forward -a in argv, it sets the global initKmerSetSize (main.c:202) -- handed on as "-a <n>";
 *   out:  overlaplen (after the odd / 13..63|127 fix-ups), thrd_num, repsTie, deLowKmer stay set; initKmerSetSize is reset to 0
 *         (pregraph.c:122).  The globals are the reference's own (inc/global.h:28,67,71,79,90); they are declared weak so that the
 *         shim also links where they do not exist.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/pregraph_b200.h"
#ifndef PGB_FLAVOUR127
#error "compile with -DPGB_FLAVOUR127=0 (SOAPdenovo-63mer) or -DPGB_FLAVOUR127=1 (SOAPdenovo-127mer)"
#endif
extern int initKmerSetSize __attribute__((weak));
extern int overlaplen __attribute__((weak));
extern int thrd_num __attribute__((weak));
extern char repsTie __attribute__((weak));     /* `boolean` is a char, inc/def2.h:25 */
extern char deLowKmer __attribute__((weak));

int call_pregraph(int argc, char **argv)
{
    char abuf[16];
    char *av[64];
    int n = 0, i, have_a = 0, K = 23, P = 8, R = 0, D = 0, rc;
    for (i = 0; i < argc && n < 60; i++) {
        av[n++] = argv[i];
        if (strncmp(argv[i], "-a", 2) == 0) have_a = 1;
        else if (strcmp(argv[i], "-R") == 0) R = 1;
        else if (i + 1 < argc && strcmp(argv[i], "-K") == 0) K = atoi(argv[i + 1]);
        else if (i + 1 < argc && strcmp(argv[i], "-p") == 0) P = atoi(argv[i + 1]);
        else if (i + 1 < argc && strcmp(argv[i], "-d") == 0) D = atoi(argv[i + 1]) >= 0 ? atoi(argv[i + 1]) : 0;
    }
    if (&initKmerSetSize && initKmerSetSize > 0 && !have_a) {
        snprintf(abuf, sizeof abuf, "%d", initKmerSetSize);
        av[n++] = "-a";
        av[n++] = abuf;
    }
    av[n] = NULL;
    rc = pgb200_pregraph_main(n, av, PGB_FLAVOUR127);
    if (K % 2 == 0) K++;
    if (K < 13) K = 13;
    if (K > (PGB_FLAVOUR127 ? 127 : 63)) K = PGB_FLAVOUR127 ? 127 : 63;
    if (&overlaplen) overlaplen = K;
    if (&thrd_num) thrd_num = P;
    if (&repsTie) repsTie = (char)R;
    if (&deLowKmer) deLowKmer = (char)D;
    if (&initKmerSetSize) initKmerSetSize = 0;
    return rc;
}
