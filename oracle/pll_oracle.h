/*
 * pll_oracle.h - prototypes of the CPU oracle (TEST INFRASTRUCTURE ONLY, see pll_oracle.c).
 * Same argument lists as the reference's pll_core_* functions (reference src/pll.h:829-1027,
 * 1659-1672), except orc_core_update_partial_tt which takes the two P-matrices instead of the
 * reference's pre-multiplied lookup table.
 */
#ifndef PLL_ORACLE_H_
#define PLL_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

int orc_core_update_pmatrix(double ** pmatrix, unsigned int states, unsigned int rate_cats,
                            const double * rates, const double * branch_lengths,
                            const unsigned int * matrix_indices,
                            const unsigned int * params_indices, const double * prop_invar,
                            double * const * eigenvals, double * const * eigenvecs,
                            double * const * inv_eigenvecs, unsigned int count,
                            unsigned int attrib);

void orc_core_update_partial_ii(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                double * parent_clv, unsigned int * parent_scaler,
                                const double * left_clv, const double * right_clv,
                                const double * left_matrix, const double * right_matrix,
                                const unsigned int * left_scaler, const unsigned int * right_scaler,
                                unsigned int attrib);

void orc_core_update_partial_ti(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                double * parent_clv, unsigned int * parent_scaler,
                                const unsigned char * left_tipchars, const double * right_clv,
                                const double * left_matrix, const double * right_matrix,
                                const unsigned int * right_scaler, const unsigned int * tipmap,
                                unsigned int tipmap_size, unsigned int attrib);

void orc_core_update_partial_tt(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                double * parent_clv, unsigned int * parent_scaler,
                                const unsigned char * left_tipchars,
                                const unsigned char * right_tipchars, const double * left_matrix,
                                const double * right_matrix, const unsigned int * tipmap,
                                unsigned int tipmap_size, unsigned int attrib);

double orc_core_edge_loglikelihood_ii(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                      const double * parent_clv, const unsigned int * parent_scaler,
                                      const double * child_clv, const unsigned int * child_scaler,
                                      const double * pmatrix, double * const * frequencies,
                                      const double * rate_weights,
                                      const unsigned int * pattern_weights,
                                      const double * invar_proportion, const int * invar_indices,
                                      const unsigned int * freqs_indices, double * persite_lnl,
                                      unsigned int attrib);

double orc_core_edge_loglikelihood_ti(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                      const double * parent_clv, const unsigned int * parent_scaler,
                                      const unsigned char * tipchars, const unsigned int * tipmap,
                                      unsigned int tipmap_size, const double * pmatrix,
                                      double * const * frequencies, const double * rate_weights,
                                      const unsigned int * pattern_weights,
                                      const double * invar_proportion, const int * invar_indices,
                                      const unsigned int * freqs_indices, double * persite_lnl,
                                      unsigned int attrib);

double orc_core_root_loglikelihood(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                   const double * clv, const unsigned int * scaler,
                                   double * const * frequencies, const double * rate_weights,
                                   const unsigned int * pattern_weights,
                                   const double * invar_proportion, const int * invar_indices,
                                   const unsigned int * freqs_indices, double * persite_lnl,
                                   unsigned int attrib);

int orc_core_update_sumtable_ii(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                const double * parent_clv, const double * child_clv,
                                const unsigned int * parent_scaler, const unsigned int * child_scaler,
                                double * const * eigenvecs, double * const * inv_eigenvecs,
                                double * const * freqs, double * sumtable, unsigned int attrib);

int orc_core_update_sumtable_ti(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                const double * parent_clv, const unsigned char * left_tipchars,
                                const unsigned int * parent_scaler, double * const * eigenvecs,
                                double * const * inv_eigenvecs, double * const * freqs,
                                const unsigned int * tipmap, unsigned int tipmap_size,
                                double * sumtable, unsigned int attrib);

int orc_core_likelihood_derivatives(unsigned int states, unsigned int sites, unsigned int rate_cats,
                                    const double * rate_weights, const unsigned int * parent_scaler,
                                    const unsigned int * child_scaler, const int * invariant,
                                    const unsigned int * pattern_weights, double branch_length,
                                    const double * prop_invar, double * const * freqs,
                                    const double * rates, double * const * eigenvals,
                                    const double * sumtable, double * d_f, double * dd_f,
                                    unsigned int attrib);

#ifdef __cplusplus
}
#endif
#endif
