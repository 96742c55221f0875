/* mini_cjson.c -- a small JSON reader/writer exposing the subset of the cJSON API the reference's
 * handlers call (numbers, strings, arrays, objects; linked children exactly like cJSON, so
 * cJSON_GetArrayItem is the same O(n) walk).  TEST INFRASTRUCTURE: both the reference build and
 * the drop-in build of tests/test_handlers_e2e.py use this same file, so its number formatting
 * only has to be deterministic, not cJSON's. */
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cjson/cJSON.h"

enum { T_NULL = 1, T_BOOL, T_NUMBER, T_STRING, T_ARRAY, T_OBJECT };
static const char *g_err;

static cJSON *node(int type) {
    cJSON *n = (cJSON *)calloc(1, sizeof *n);
    n->type = type;
    return n;
}
void cJSON_Delete(cJSON *i) {
    while (i) {
        cJSON *nx = i->next;
        cJSON_Delete(i->child);
        free(i->valuestring);
        free(i->string);
        free(i);
        i = nx;
    }
}
static const char *skip(const char *p) { while (*p && isspace((unsigned char)*p)) p++; return p; }
static const char *parse_value(const char *p, cJSON **out);
static const char *parse_string_raw(const char *p, char **out) {
    if (*p != '"') return NULL;
    const char *q = ++p;
    while (*q && *q != '"') q += (*q == '\\' && q[1]) ? 2 : 1;
    if (*q != '"') return NULL;
    char *s = (char *)malloc((size_t)(q - p) + 1), *d = s;
    while (p < q) {
        if (*p == '\\') { p++; *d++ = *p == 'n' ? '\n' : *p == 't' ? '\t' : *p; p++; }
        else *d++ = *p++;
    }
    *d = 0;
    *out = s;
    return q + 1;
}
static const char *parse_value(const char *p, cJSON **out) {
    p = skip(p);
    if (*p == '"') {
        cJSON *n = node(T_STRING);
        p = parse_string_raw(p, &n->valuestring);
        if (!p) { cJSON_Delete(n); return NULL; }
        *out = n;
        return p;
    }
    if (*p == '[' || *p == '{') {
        const int obj = *p == '{';
        cJSON *n = node(obj ? T_OBJECT : T_ARRAY), *tail = NULL;
        p = skip(p + 1);
        if (*p == (obj ? '}' : ']')) { *out = n; return p + 1; }
        for (;;) {
            char *key = NULL;
            if (obj) {
                p = parse_string_raw(skip(p), &key);
                if (!p) { cJSON_Delete(n); return NULL; }
                p = skip(p);
                if (*p != ':') { free(key); cJSON_Delete(n); return NULL; }
                p++;
            }
            cJSON *c = NULL;
            p = parse_value(p, &c);
            if (!p) { free(key); cJSON_Delete(n); return NULL; }
            c->string = key;
            if (tail) { tail->next = c; c->prev = tail; } else n->child = c;
            tail = c;
            p = skip(p);
            if (*p == ',') { p++; continue; }
            if (*p == (obj ? '}' : ']')) { *out = n; return p + 1; }
            cJSON_Delete(n);
            return NULL;
        }
    }
    if (!strncmp(p, "true", 4) || !strncmp(p, "false", 5)) {
        cJSON *n = node(T_BOOL);
        n->valueint = *p == 't';
        *out = n;
        return p + (*p == 't' ? 4 : 5);
    }
    if (!strncmp(p, "null", 4)) { *out = node(T_NULL); return p + 4; }
    char *end;
    const double v = strtod(p, &end);
    if (end == p) return NULL;
    cJSON *n = node(T_NUMBER);
    n->valuedouble = v;
    n->valueint = (int)v;
    *out = n;
    return end;
}
cJSON *cJSON_Parse(const char *value) {
    cJSON *n = NULL;
    const char *p = parse_value(value, &n);
    if (!p) { g_err = value; return NULL; }
    return n;
}
const char *cJSON_GetErrorPtr(void) { return g_err; }
cJSON *cJSON_GetObjectItem(const cJSON *o, const char *s) {
    for (cJSON *c = o ? o->child : NULL; c; c = c->next)
        if (c->string && strcmp(c->string, s) == 0) return c;
    return NULL;
}
int cJSON_GetArraySize(const cJSON *a) {
    int n = 0;
    for (cJSON *c = a ? a->child : NULL; c; c = c->next) n++;
    return n;
}
cJSON *cJSON_GetArrayItem(const cJSON *a, int index) {
    cJSON *c = a ? a->child : NULL;
    while (c && index-- > 0) c = c->next;
    return c;
}
cJSON_bool cJSON_IsNumber(const cJSON *i) { return i && i->type == T_NUMBER; }
cJSON_bool cJSON_IsString(const cJSON *i) { return i && i->type == T_STRING; }
cJSON_bool cJSON_IsArray(const cJSON *i) { return i && i->type == T_ARRAY; }
cJSON *cJSON_CreateObject(void) { return node(T_OBJECT); }
cJSON *cJSON_CreateArray(void) { return node(T_ARRAY); }
cJSON *cJSON_CreateNumber(double num) {
    cJSON *n = node(T_NUMBER);
    n->valuedouble = num;
    n->valueint = (int)num;
    return n;
}
static void append(cJSON *parent, cJSON *item) {
    cJSON *c = parent->child;
    if (!c) { parent->child = item; return; }
    while (c->next) c = c->next;
    c->next = item;
    item->prev = c;
}
cJSON_bool cJSON_AddItemToArray(cJSON *a, cJSON *i) { append(a, i); return 1; }
cJSON_bool cJSON_AddItemToObject(cJSON *o, const char *s, cJSON *i) {
    free(i->string);
    i->string = strdup(s);
    append(o, i);
    return 1;
}
cJSON *cJSON_CreateDoubleArray(const double *numbers, int count) {
    cJSON *a = cJSON_CreateArray(), *tail = NULL;
    for (int i = 0; i < count; i++) {
        cJSON *n = cJSON_CreateNumber(numbers[i]);
        if (tail) { tail->next = n; n->prev = tail; } else a->child = n;
        tail = n;
    }
    return a;
}
cJSON *cJSON_AddNumberToObject(cJSON *o, const char *name, double number) {
    cJSON *n = cJSON_CreateNumber(number);
    cJSON_AddItemToObject(o, name, n);
    return n;
}
cJSON *cJSON_AddStringToObject(cJSON *o, const char *name, const char *string) {
    cJSON *n = node(T_STRING);
    n->valuestring = strdup(string);
    cJSON_AddItemToObject(o, name, n);
    return n;
}

typedef struct { char *p; size_t n, cap; } sbuf;
static void put(sbuf *b, const char *s) {
    const size_t l = strlen(s);
    if (b->n + l + 1 > b->cap) { b->cap = (b->n + l + 1) * 2; b->p = (char *)realloc(b->p, b->cap); }
    memcpy(b->p + b->n, s, l + 1);
    b->n += l;
}
static void print_value(sbuf *b, const cJSON *i) {
    char tmp[64];
    switch (i->type) {
        case T_NULL: put(b, "null"); break;
        case T_BOOL: put(b, i->valueint ? "true" : "false"); break;
        case T_NUMBER:
            if (isnan(i->valuedouble) || isinf(i->valuedouble)) { put(b, "null"); break; }
            snprintf(tmp, sizeof tmp, "%1.15g", i->valuedouble);
            if (strtod(tmp, NULL) != i->valuedouble) snprintf(tmp, sizeof tmp, "%1.17g", i->valuedouble);
            put(b, tmp);
            break;
        case T_STRING: put(b, "\""); put(b, i->valuestring ? i->valuestring : ""); put(b, "\""); break;
        default: {
            const int obj = i->type == T_OBJECT;
            put(b, obj ? "{" : "[");
            for (cJSON *c = i->child; c; c = c->next) {
                if (obj) { put(b, "\""); put(b, c->string ? c->string : ""); put(b, "\":"); }
                print_value(b, c);
                if (c->next) put(b, ",");
            }
            put(b, obj ? "}" : "]");
        }
    }
}
char *cJSON_PrintUnformatted(const cJSON *item) {
    sbuf b = {NULL, 0, 0};
    put(&b, "");
    print_value(&b, item);
    return b.p;
}
