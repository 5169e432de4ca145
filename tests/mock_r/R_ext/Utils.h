/* R_ext/Utils.h (MOCK) -- see ../Rinternals.h. */
#ifndef MOCK_R_UTILS_H
#define MOCK_R_UTILS_H
#ifdef __cplusplus
extern "C" {
#endif
void R_CheckUserInterrupt(void);
#ifdef __cplusplus
}
#endif
#endif
