/* input.h -- "key = value" input files with the reference's semantics (parse_input.c:30-104):
 * '#' comments, whitespace trimmed, first match wins, an empty value counts as absent,
 * unknown keys are ignored.  Key names are the reference's (README.md:18-80). */
#ifndef PPH_INPUT_H_
#define PPH_INPUT_H_

typedef struct pph_input pph_input;

pph_input *pph_input_load(const char *path);          /* NULL if the file cannot be opened */
void pph_input_free(pph_input *in);

/* return 1 if the key was found, 0 otherwise; a missing mandatory key is fatal
 * ("Mandatory key '%s' not found, exiting", exit status 255 like exit(-1), parse_input.c:97-100) */
int pph_get_string(pph_input *in, const char *key, char *dest, int dest_len, int mandatory);
int pph_get_int(pph_input *in, const char *key, int *dest, int mandatory);
int pph_get_llint(pph_input *in, const char *key, long long *dest, int mandatory);
int pph_get_double(pph_input *in, const char *key, double *dest, int mandatory);

#endif
