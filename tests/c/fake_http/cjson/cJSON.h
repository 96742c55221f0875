/* Minimal stand-in for <cjson/cJSON.h>; see microhttpd.h in this directory. */
#ifndef SVDB_TEST_STUB_CJSON_H
#define SVDB_TEST_STUB_CJSON_H
typedef struct cJSON {
    struct cJSON *next, *prev, *child;
    int type;
    char *valuestring;
    int valueint;
    double valuedouble;
    char *string;
} cJSON;
typedef int cJSON_bool;
cJSON *cJSON_Parse(const char *value);
void cJSON_Delete(cJSON *item);
const char *cJSON_GetErrorPtr(void);
cJSON *cJSON_GetObjectItem(const cJSON *object, const char *string);
int cJSON_GetArraySize(const cJSON *array);
cJSON *cJSON_GetArrayItem(const cJSON *array, int index);
cJSON_bool cJSON_IsNumber(const cJSON *item);
cJSON_bool cJSON_IsString(const cJSON *item);
cJSON_bool cJSON_IsArray(const cJSON *item);
cJSON *cJSON_CreateObject(void);
cJSON *cJSON_CreateArray(void);
cJSON *cJSON_CreateNumber(double num);
cJSON *cJSON_CreateDoubleArray(const double *numbers, int count);
cJSON *cJSON_AddNumberToObject(cJSON *object, const char *name, double number);
cJSON *cJSON_AddStringToObject(cJSON *object, const char *name, const char *string);
cJSON_bool cJSON_AddItemToObject(cJSON *object, const char *string, cJSON *item);
cJSON_bool cJSON_AddItemToArray(cJSON *array, cJSON *item);
char *cJSON_PrintUnformatted(const cJSON *item);
#endif
