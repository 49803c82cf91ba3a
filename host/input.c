#define _GNU_SOURCE
#include "input.h"

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
	char *key, *value;
} pph_pair;

struct pph_input {
	pph_pair *pairs;
	int n, cap;
};

static char *trimmed_copy(const char *b, const char *e) {
	while(b < e && isspace((unsigned char) *b)) b++;
	while(e > b && isspace((unsigned char) e[-1])) e--;
	char *out = malloc((size_t) (e - b) + 1);
	memcpy(out, b, (size_t) (e - b));
	out[e - b] = '\0';
	return out;
}

pph_input *pph_input_load(const char *path) {
	FILE *f = fopen(path, "r");
	if(f == NULL) return NULL;
	pph_input *in = calloc(1, sizeof(*in));
	char *line = NULL;
	size_t cap = 0;
	while(getline(&line, &cap, f) != -1) {
		const char *b = line, *e = line + strlen(line);
		while(b < e && isspace((unsigned char) *b)) b++;
		if(b == e || *b == '#') continue;
		const char *eq = strchr(b, '=');
		if(eq == NULL) {
			fprintf(stderr, "WARNING: malformed line '%.*s'\n", (int) (e - b - (e[-1] == '\n')), b);
			continue;
		}
		if(in->n == in->cap) {
			in->cap += 16;
			in->pairs = realloc(in->pairs, sizeof(pph_pair) * (size_t) in->cap);
		}
		in->pairs[in->n].key = trimmed_copy(b, eq);
		in->pairs[in->n].value = trimmed_copy(eq + 1, e);
		in->n++;
	}
	free(line);
	fclose(f);
	return in;
}

void pph_input_free(pph_input *in) {
	if(in == NULL) return;
	for(int i = 0; i < in->n; i++) {
		free(in->pairs[i].key);
		free(in->pairs[i].value);
	}
	free(in->pairs);
	free(in);
}

static const char *lookup(pph_input *in, const char *key, int mandatory) {
	for(int i = 0; i < in->n; i++)
		if(strcmp(in->pairs[i].key, key) == 0 && in->pairs[i].value[0] != '\0') return in->pairs[i].value;
	if(mandatory) {
		fprintf(stderr, "Mandatory key '%s' not found, exiting\n", key);
		exit(-1);
	}
	fprintf(stderr, "INFO: optional key '%s' not found\n", key);
	return NULL;
}

int pph_get_string(pph_input *in, const char *key, char *dest, int dest_len, int mandatory) {
	const char *v = lookup(in, key, mandatory);
	if(v == NULL) return 0;
	snprintf(dest, (size_t) dest_len, "%s", v);
	return 1;
}

int pph_get_int(pph_input *in, const char *key, int *dest, int mandatory) {
	const char *v = lookup(in, key, mandatory);
	if(v == NULL) return 0;
	*dest = atoi(v);
	return 1;
}

int pph_get_llint(pph_input *in, const char *key, long long *dest, int mandatory) {
	const char *v = lookup(in, key, mandatory);
	if(v == NULL) return 0;
	*dest = atoll(v);
	return 1;
}

int pph_get_double(pph_input *in, const char *key, double *dest, int mandatory) {
	const char *v = lookup(in, key, mandatory);
	if(v == NULL) return 0;
	*dest = atof(v);
	return 1;
}
