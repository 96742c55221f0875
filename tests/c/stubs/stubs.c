/* Link-time stand-ins for libmicrohttpd and libcjson (never executed; see microhttpd.h here). */
#include <stdlib.h>
#include "microhttpd.h"
#include "cjson/cJSON.h"
struct MHD_Response *MHD_create_response_from_buffer(size_t s, void *b, enum MHD_ResponseMemoryMode m) { (void)s; (void)b; (void)m; return NULL; }
enum MHD_Result MHD_add_response_header(struct MHD_Response *r, const char *h, const char *c) { (void)r; (void)h; (void)c; return MHD_NO; }
enum MHD_Result MHD_queue_response(struct MHD_Connection *c, unsigned int s, struct MHD_Response *r) { (void)c; (void)s; (void)r; return MHD_NO; }
void MHD_destroy_response(struct MHD_Response *r) { (void)r; }
const char *MHD_lookup_connection_value(struct MHD_Connection *c, enum MHD_ValueKind k, const char *key) { (void)c; (void)k; (void)key; return NULL; }
struct MHD_Daemon *MHD_start_daemon(unsigned int f, uint16_t p, MHD_AcceptPolicyCallback a, void *ac, MHD_AccessHandlerCallback d, void *dc, ...) { (void)f; (void)p; (void)a; (void)ac; (void)d; (void)dc; return NULL; }
void MHD_stop_daemon(struct MHD_Daemon *d) { (void)d; }
cJSON *cJSON_Parse(const char *v) { (void)v; return NULL; }
void cJSON_Delete(cJSON *i) { (void)i; }
const char *cJSON_GetErrorPtr(void) { return NULL; }
cJSON *cJSON_GetObjectItem(const cJSON *o, const char *s) { (void)o; (void)s; return NULL; }
int cJSON_GetArraySize(const cJSON *a) { (void)a; return 0; }
cJSON *cJSON_GetArrayItem(const cJSON *a, int i) { (void)a; (void)i; return NULL; }
cJSON_bool cJSON_IsNumber(const cJSON *i) { (void)i; return 0; }
cJSON_bool cJSON_IsString(const cJSON *i) { (void)i; return 0; }
cJSON_bool cJSON_IsArray(const cJSON *i) { (void)i; return 0; }
cJSON *cJSON_CreateObject(void) { return NULL; }
cJSON *cJSON_CreateArray(void) { return NULL; }
cJSON *cJSON_CreateNumber(double n) { (void)n; return NULL; }
cJSON *cJSON_CreateDoubleArray(const double *n, int c) { (void)n; (void)c; return NULL; }
cJSON *cJSON_AddNumberToObject(cJSON *o, const char *n, double v) { (void)o; (void)n; (void)v; return NULL; }
cJSON *cJSON_AddStringToObject(cJSON *o, const char *n, const char *s) { (void)o; (void)n; (void)s; return NULL; }
cJSON_bool cJSON_AddItemToObject(cJSON *o, const char *s, cJSON *i) { (void)o; (void)s; (void)i; return 0; }
cJSON_bool cJSON_AddItemToArray(cJSON *a, cJSON *i) { (void)a; (void)i; return 0; }
char *cJSON_PrintUnformatted(const cJSON *i) { (void)i; return NULL; }
