/* R_ext/Rallocators.h (MOCK): custom allocators for vectors ("Writing R Extensions" 6.1.1 / R >= 3.1.0): R calls
 * mem_alloc(allocator, header + data bytes) instead of malloc and mem_free(allocator, p) when the object is collected. */
#ifndef MOCK_RALLOCATORS_H
#define MOCK_RALLOCATORS_H
#include <stddef.h>
#include <Rinternals.h>
typedef struct R_allocator R_allocator_t;
typedef void *(*custom_alloc_t)(R_allocator_t *allocator, size_t);
typedef void (*custom_free_t)(R_allocator_t *allocator, void *);
struct R_allocator {
  custom_alloc_t mem_alloc;
  custom_free_t mem_free;
  void *res;
  void *data;
};
#ifdef __cplusplus
extern "C" {
#endif
SEXP Rf_allocVector3(SEXPTYPE, R_xlen_t, R_allocator_t *);
int mock_custom_allocations(void);   /* test hook: vectors currently living in allocator-provided memory */
#ifdef __cplusplus
}
#endif
#endif
