/* fake_mhd.c -- an in-process stand-in for the handful of libmicrohttpd calls the reference's
 * handlers make.  A "connection" is a struct holding the query-string arguments of the request
 * and capturing the queued response.  TEST INFRASTRUCTURE (tests/test_handlers_e2e.py). */
#include <stdlib.h>
#include <string.h>

#include "fake_mhd.h"

struct MHD_Response {
    char *body;
    size_t size;
};

struct MHD_Response *MHD_create_response_from_buffer(size_t size, void *buffer, enum MHD_ResponseMemoryMode mode) {
    struct MHD_Response *r = (struct MHD_Response *)calloc(1, sizeof *r);
    r->body = (char *)malloc(size + 1);
    memcpy(r->body, buffer, size);
    r->body[size] = 0;
    r->size = size;
    if (mode == MHD_RESPMEM_MUST_FREE) free(buffer);
    return r;
}
enum MHD_Result MHD_add_response_header(struct MHD_Response *r, const char *h, const char *c) {
    (void)r; (void)h; (void)c;
    return MHD_YES;
}
enum MHD_Result MHD_queue_response(struct MHD_Connection *c, unsigned int status, struct MHD_Response *r) {
    if (!c || !r) return MHD_NO;
    free(c->body);
    c->body = (char *)malloc(r->size + 1);
    memcpy(c->body, r->body, r->size + 1);
    c->status = status;
    c->responded = 1;
    return MHD_YES;
}
void MHD_destroy_response(struct MHD_Response *r) {
    if (!r) return;
    free(r->body);
    free(r);
}
const char *MHD_lookup_connection_value(struct MHD_Connection *c, enum MHD_ValueKind kind, const char *key) {
    if (kind != MHD_GET_ARGUMENT_KIND) return NULL;
    for (int i = 0; i < c->nargs; i++)
        if (strcmp(c->keys[i], key) == 0) return c->vals[i];
    return NULL;
}
struct MHD_Daemon *MHD_start_daemon(unsigned int f, uint16_t p, MHD_AcceptPolicyCallback a, void *ac,
                                    MHD_AccessHandlerCallback d, void *dc, ...) {
    (void)f; (void)p; (void)a; (void)ac; (void)d; (void)dc;
    return NULL;
}
void MHD_stop_daemon(struct MHD_Daemon *d) { (void)d; }
