# Optional replacement for the two grid loops of tune() (reference R/insider.R:100-132, 145-174): identical semantics
# (two phases, round(.,2), which.min on test RMSE, CSV side effects), but the data and masks are uploaded to the GPU once.
tune_b200 <- function(object, latent_dimension = NULL, lambda = 0.1, alpha = 0.0) {
    if(!is.integer(latent_dimension) | !is.numeric(lambda) | !is.numeric(alpha)){
        stop("TUNNING: The element of latent_dimension, lambda, and alpha should be integer, numeric, and numeric.")
    }
    if((length(latent_dimension) <= 1) & (length(lambda) <= 1 & length(alpha) <= 1)){
        stop("TUNNING: The length of either latent_dimension or lambda and alpha should be greater than 1.")
    }
    p <- object[['params']]
    h <- b200_upload(object[['data']], object[['confounder']], object[['ctns_confounder']], object[['train_indicator']],
                     object[['test_indicator']], object[['inc_continuous']])
    on.exit(b200_release(h))
    init <- function(latent_rank) {
        confounder_num <- ncol(object[['confounder']])
        l <- lapply(1:confounder_num, function(i) matrix(init_parameters(length(unique(object[['confounder']][,i])) * latent_rank), ncol = latent_rank))
        if(object[['inc_continuous']] == 1) l[[confounder_num + 1]] <- matrix(init_parameters(ncol(object[['ctns_confounder']]) * latent_rank), ncol = latent_rank)
        list(l, matrix(init_parameters(latent_rank * ncol(object[['data']])), nrow = latent_rank))
    }
    rank_tuning <- NULL; reg_tuning <- NULL
    if(length(latent_dimension) > 1){
        for(latent_rank in latent_dimension){
            cat('Latent rank: ', latent_rank, "---------------------------------\n")
            f <- init(latent_rank)
            hp <- if(length(lambda) == 1 & length(alpha) == 1) c(lambda, lambda, alpha) else c(0.1, 0.1, 0)
            fitted <- b200_optimize_resident(h, f[[1]], f[[2]], latent_rank, hp[1], hp[2], hp[3], 1, p[['global_tol']], p[['sub_tol']], p[['tuning_iter']])
            rank_tuning <- rbind(rank_tuning, c(latent_rank, fitted$train_rmse, fitted$test_rmse))
            write.csv(rank_tuning, file = 'insider_rank_tuning_result.csv')
        }
    }
    latent_rank <- if(length(latent_dimension) > 1) latent_dimension[which.min(rank_tuning[,3])] else latent_dimension
    if(length(lambda) > 1 | length(alpha) > 1){
        param_grid <- expand.grid(lambda = lambda, alpha = alpha)
        for(i in seq(nrow(param_grid))){
            cat('parameter grid:', paste(round(param_grid[i,], 2), collapse = ','), "---------------------------------\n")
            l <- round(param_grid[i, 1], 2); a <- round(param_grid[i, 2], 2)
            f <- init(latent_rank)
            fitted <- b200_optimize_resident(h, f[[1]], f[[2]], latent_rank, l, l, a, 1, p[['global_tol']], p[['sub_tol']], p[['tuning_iter']])
            reg_tuning <- rbind(reg_tuning, c(l, a, fitted$train_rmse, fitted$test_rmse))
            write.csv(reg_tuning, file = paste0('insider_R', latent_rank, '_reg_tuning_result.csv'))
        }
    }
    list(rank_tuning = rank_tuning, latent_rank = latent_rank, reg_tuning = reg_tuning)
}
