// TEST INFRASTRUCTURE (oracle/_ref build). The reference's compile-time configuration header is withheld
// (.gitignore:18 of the reference); these are the ten macros its open sources consume, with
// the values the survey's probe validated on all goldens. The reduced Groebner basis does not depend on them.
#pragma once
#define CACHE_LINE_SIZE 64
#define HASH_LOAD_FACTOR 0.5
#define HASHTABLE_INIT_SIZE (1UL << 16)
#define SPAIRSET_INIT_SIZE (1UL << 12)
#define NUM_ROWS_BUFFER 64
#define USE_SPARSEST_REDUCER 1
#define USE_DELETED_REDUCERS 0
#define INSERT_ELEMENTS_DECREASING 1
#define NONDEG_ORDERS_HEURISTIC 1
#define GAMBA_VERSION "0.1.0"
