/*
 * Force-included (gcc -include) ahead of the reference's sources by
 * oracle/Makefile.  Two jobs, neither touches the reference files:
 *  1. <stdint.h>: src/vector_database.c:86-90 uses SIZE_MAX without including
 *     it (builds on macOS by transitive include, fails on glibc).
 *  2. printf -> no-op: kdtree.c:16,48,89 and vector_database.c:84-89 print one
 *     or more lines per insert level; formatting them dominates build time.
 *     stderr diagnostics (fprintf) are left alone.  Arithmetic is unaffected.
 */
#ifndef SVDB_ORACLE_QUIET_STDIO_H
#define SVDB_ORACLE_QUIET_STDIO_H
#include <stdint.h>
#include <stdio.h>
#define printf(...) ((void)0)
#endif
