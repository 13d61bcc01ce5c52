# glm_interaction() with the reference's signature and return value (R/glm_interaction.R:2-30): list(coeff_matrix, pval_matrix),
# row i = interaction level i. The per-level regressions run on the GPU (insider_b200_glm_interaction); train_indicator, tol and
# n_cores are accepted and unused, as in the reference.
#' @export
glm_interaction <- function(residual, train_indicator, interaction_indicator, column_factor, tol = 1e-10, n_cores = 10){
    b200_glm_interaction(residual, as.integer(interaction_indicator), column_factor)
}
