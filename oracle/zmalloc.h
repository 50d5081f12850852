/* TEST INFRASTRUCTURE (oracle).  Forced-include for the "z" build of the
 * reference: every malloc in the reference sources becomes a zeroing calloc so
 * that drivers which read a work vector before writing it (GPBiCG/GPBiCR `mr`,
 * SURVEY.md App. B.11) give run-to-run deterministic results. */
#ifndef ORACLE_ZMALLOC_H
#define ORACLE_ZMALLOC_H
#include <stdlib.h>
#define malloc(x) calloc(1, (x))
#endif
