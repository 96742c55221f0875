/* handler_driver.c -- replays a script of HTTP requests through the REFERENCE'S OWN handler
 * functions (src/*_handler.c, unmodified) without a network: routing as in src/main.c:241-279,
 * call protocol as libmicrohttpd drives it (first call with *con_cls == NULL, then the upload
 * chunk, then a call with *upload_data_size == 0).  Prints "<status> <body>" per request.
 * Linked twice: with the reference's vector_database.c + kdtree.c, and with libsvdb_b200.so.
 *
 * usage: driver <script> <kd_dim> <vector_size>
 * script lines:  METHOD URL key=val&key=val BODY...   ("-" for no args / no body)
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fake_mhd.h"
#include "compare_handler.h"
#include "delete_handler.h"
#include "get_handler.h"
#include "post_handler.h"
#include "put_handler.h"

static MHD_AccessHandlerCallback route(const char *method, const char *url) {
    if (!strcmp(method, "GET")) {
        if (!strcmp(url, "/vector")) return get_handler;
        if (!strncmp(url, "/compare/", 9)) return compare_handler;
    } else if (!strcmp(method, "POST")) {
        if (!strcmp(url, "/vector")) return post_handler;
        if (!strcmp(url, "/nearest")) return nearest_handler;
#ifdef SVDB_PATCHED_HANDLERS                                 /* integration/f3_f4_handlers.patch: the route it adds to main.c */
        if (!strncmp(url, "/compare/", 9)) return compare_batch_handler;
#endif
    } else if (!strcmp(method, "PUT") && !strcmp(url, "/vector")) {
        return put_handler;
    } else if (!strcmp(method, "DELETE") && !strcmp(url, "/vector")) {
        return delete_handler;
    }
    return NULL;
}

static void parse_args(struct MHD_Connection *c, char *args) {
    c->nargs = 0;
    if (!strcmp(args, "-")) return;
    for (char *tok = strtok(args, "&"); tok && c->nargs < 8; tok = strtok(NULL, "&")) {
        char *eq = strchr(tok, '=');
        if (!eq) continue;
        *eq = 0;
        snprintf(c->keys[c->nargs], sizeof c->keys[0], "%s", tok);
        snprintf(c->vals[c->nargs], sizeof c->vals[0], "%s", eq + 1);
        c->nargs++;
    }
}

int main(int argc, char **argv) {
    if (argc < 4) return 64;
    FILE *f = fopen(argv[1], "r");
    if (!f) return 65;
    PostHandlerData data;
    data.db = vector_db_init(0, (size_t)atol(argv[2]));     /* src/main.c:351 */
    data.db_vector_size = (size_t)atol(argv[3]);
    if (!data.db) return 66;
    if (!freopen("/dev/null", "w", stderr)) return 67;       /* the handlers chat on stderr */
    FILE *out = fdopen(dup(1), "w");
    if (!freopen("/dev/null", "w", stdout)) return 68;       /* ... and on stdout */

    size_t cap = 1 << 22;
    char *line = (char *)malloc(cap);
    long n = 0;
    while (fgets(line, (int)cap, f)) {
        line[strcspn(line, "\n")] = 0;
        char *method = strtok(line, " "), *url = strtok(NULL, " "), *args = strtok(NULL, " "), *body = strtok(NULL, "");
        if (!method || !url || !args) continue;
        struct MHD_Connection conn;
        memset(&conn, 0, sizeof conn);
        char argbuf[1024];
        snprintf(argbuf, sizeof argbuf, "%s", args);
        parse_args(&conn, argbuf);
        MHD_AccessHandlerCallback h = route(method, url);
        if (!h) {
            fprintf(out, "%ld 404 no route\n", n++);
            continue;
        }
        void *con_cls = NULL;
        size_t size = 0;
        h(&data, &conn, url, method, "HTTP/1.1", NULL, &size, &con_cls);
        if (!conn.responded && body && strcmp(body, "-")) {
            size = strlen(body);
            h(&data, &conn, url, method, "HTTP/1.1", body, &size, &con_cls);
        }
        if (!conn.responded) {
            size = 0;
            h(&data, &conn, url, method, "HTTP/1.1", NULL, &size, &con_cls);
        }
        fprintf(out, "%ld %u %s\n", n++, conn.status, conn.responded ? conn.body : "(no response)");
        free(conn.body);
        if (con_cls) {                                       /* src/main.c:289-297 request_completed_callback */
            struct { char *data; size_t data_size; } *cd = con_cls;
            free(cd->data);
            free(cd);
        }
    }
    fprintf(out, "final size=%zu\n", data.db->size);
    fflush(out);
    vector_db_free(data.db);
    fclose(f);
    return 0;
}
