// ga.cu -- genetic-algorithm entry points (a15 / a16). PLACEHOLDER: filled in after the LM path is measured.
#include "common.cuh"
using namespace pnol;
#define PNOL_GA_TODO(ctx) do { if (ctx) PNOL_SET_ERR(ctx, "GA path not built yet"); return PNOL_ERR_NO_FUNCTOR; } while (0)
extern "C" int pnol_ga_create(pnol_ctx * ctx, const pnol_functor *, const pnol_ga_params *, int, const double *, const double *, const pnol_stream_desc *, pnol_ga **) { PNOL_GA_TODO(ctx); }
extern "C" void pnol_ga_destroy(pnol_ga *) {}
extern "C" int pnol_ga_init(pnol_ga *, const double *, double *) { return PNOL_ERR_NO_FUNCTOR; }
extern "C" int pnol_ga_generation(pnol_ga *) { return PNOL_ERR_NO_FUNCTOR; }
extern "C" int pnol_ga_status_get(pnol_ga *, pnol_ga_status *) { return PNOL_ERR_NO_FUNCTOR; }
extern "C" int pnol_ga_get_population(pnol_ga *, double *, double *) { return PNOL_ERR_NO_FUNCTOR; }
extern "C" int pnol_ga_get_indices(pnol_ga *, int *, int *, int *) { return PNOL_ERR_NO_FUNCTOR; }
extern "C" int pnol_ga_pop_sort(pnol_ctx * ctx, double *, double *, long long, int) { PNOL_GA_TODO(ctx); }
extern "C" int pnol_ga_check_bounds(pnol_ctx * ctx, double *, long long, int, const double *, const double *, unsigned char *, const pnol_stream_desc *, uint64_t *) { PNOL_GA_TODO(ctx); }
extern "C" int pnol_ga_check_identical(pnol_ctx * ctx, double *, long long, int, const double *, const double *, unsigned char *, const pnol_stream_desc *, uint64_t *) { PNOL_GA_TODO(ctx); }
