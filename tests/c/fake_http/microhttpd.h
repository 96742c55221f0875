/* Minimal stand-in for <microhttpd.h> (libmicrohttpd is not installed in this image).
 * ONLY for tests/test_reference_handlers_link.py: it lets the reference's unmodified
 * src/*_handler.c and src/main.c be COMPILED and LINKED against libsvdb_b200.so to prove that
 * the drop-in needs no source change.  Declarations follow libmicrohttpd's public API shapes;
 * in tests/c/fake_http these are backed by fake_mhd.c (an in-process connection object, no sockets). */
#ifndef SVDB_TEST_STUB_MICROHTTPD_H
#define SVDB_TEST_STUB_MICROHTTPD_H
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>

enum MHD_Result { MHD_NO = 0, MHD_YES = 1 };
struct MHD_Connection;
struct MHD_Response;
struct MHD_Daemon;
enum MHD_ValueKind { MHD_HEADER_KIND = 1, MHD_GET_ARGUMENT_KIND = 8 };
enum MHD_ResponseMemoryMode { MHD_RESPMEM_PERSISTENT, MHD_RESPMEM_MUST_FREE, MHD_RESPMEM_MUST_COPY };
enum MHD_RequestTerminationCode { MHD_REQUEST_TERMINATED_COMPLETED_OK = 0 };
enum MHD_FLAG { MHD_USE_THREAD_PER_CONNECTION = 4, MHD_USE_INTERNAL_POLLING_THREAD = 8 };
enum MHD_OPTION { MHD_OPTION_END = 0, MHD_OPTION_NOTIFY_COMPLETED = 4 };
#define MHD_HTTP_OK 200
#define MHD_HTTP_BAD_REQUEST 400
#define MHD_HTTP_NOT_FOUND 404
#define MHD_HTTP_INTERNAL_SERVER_ERROR 500
#define MHD_HTTP_NOT_IMPLEMENTED 501
#define MHD_HTTP_HEADER_CONTENT_TYPE "Content-Type"

typedef enum MHD_Result (*MHD_AccessHandlerCallback)(void *cls, struct MHD_Connection *connection, const char *url,
                                                     const char *method, const char *version, const char *upload_data,
                                                     size_t *upload_data_size, void **con_cls);
typedef enum MHD_Result (*MHD_AcceptPolicyCallback)(void *cls, const void *addr, unsigned addrlen);
typedef void (*MHD_RequestCompletedCallback)(void *cls, struct MHD_Connection *connection, void **con_cls,
                                             enum MHD_RequestTerminationCode toe);

struct MHD_Response *MHD_create_response_from_buffer(size_t size, void *buffer, enum MHD_ResponseMemoryMode mode);
enum MHD_Result MHD_add_response_header(struct MHD_Response *response, const char *header, const char *content);
enum MHD_Result MHD_queue_response(struct MHD_Connection *connection, unsigned int status_code, struct MHD_Response *response);
void MHD_destroy_response(struct MHD_Response *response);
const char *MHD_lookup_connection_value(struct MHD_Connection *connection, enum MHD_ValueKind kind, const char *key);
struct MHD_Daemon *MHD_start_daemon(unsigned int flags, uint16_t port, MHD_AcceptPolicyCallback apc, void *apc_cls,
                                    MHD_AccessHandlerCallback dh, void *dh_cls, ...);
void MHD_stop_daemon(struct MHD_Daemon *daemon);
#endif
