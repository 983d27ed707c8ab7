/*
 * arch-lite params.h: reader for neutral's text decks (problems/x.params) and for
 * ../arch.params / problems/neutral.tests.
 *
 * Grammar (reference problems/csp.params:1-10):
 *   name value            # trailing comment allowed
 *   name k=v k=v ...      (source / problem_<n> / "problems/x.params result=...")
 * The first whitespace-delimited token of a line is the name.
 * Call sites: main.c:29-46; neutral_data.c:24-37; omp3/neutral.c:539-545.
 */
#ifndef ARCHLITE_PARAMS_H
#define ARCHLITE_PARAMS_H

#ifdef __cplusplus
extern "C" {
#endif

#define MAX_KEYS 32
#define MAX_STR_LEN 256

/* Fatal (TERMINATE) if the name is missing. */
int get_int_parameter(const char* param_name, const char* filename);
double get_double_parameter(const char* param_name, const char* filename);

/* Fills keys (nkeys strings, MAX_STR_LEN apart) and values for the first line whose name
 * matches; returns 1 if found, 0 otherwise. */
int get_key_value_parameter(const char* param_name, const char* filename,
                            char* keys, double* values, int* nkeys);

#ifdef __cplusplus
}
#endif

#endif
