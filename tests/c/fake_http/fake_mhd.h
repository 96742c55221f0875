#ifndef SVDB_TEST_FAKE_MHD_H
#define SVDB_TEST_FAKE_MHD_H
#include "microhttpd.h"
struct MHD_Connection {
    int nargs;
    char keys[8][32];
    char vals[8][128];
    int responded;
    unsigned status;
    char *body;
};
#endif
