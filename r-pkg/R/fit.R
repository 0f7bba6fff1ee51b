# Drop-in wrappers with the signatures the reference calls (FitOCT.R:95, 110-124; priPost.R:2-16;
# ShinyInterface/server.R:341-343, 408-426).  Swap `FitOCTLib::fitExpGP` for `FitOCTb200::fitExpGP`.
# NOT RUN in this repository's image (no R): kept logic-free; all numerics are behind .Call.

.par_names <- function(kind, Nn) {
  if (kind == 0L) c(paste0("theta[", 1:3, "]"), paste0("yGP[", seq_len(Nn), "]"), "lambda", "sigma", "br", "lp__")
  else c(paste0("theta[", 1:3, "]"), "br", "lp__")
}

# Build a genuine rstan stanfit through rstan's own reader: one Stan-CSV per chain -> rstan::read_stan_csv.
.as_stanfit <- function(res, kind, Nn, chains, nb_warmup, nb_iter) {
  pn <- .par_names(kind, Nn)
  P  <- length(pn)
  sp <- c("accept_stat__", "stepsize__", "treedepth__", "n_leapfrog__", "divergent__", "energy__")
  d  <- aperm(array(res$draws, c(P, chains, nb_iter)), c(3, 2, 1))          # [iter, chain, par]
  s  <- aperm(array(res$sampler_params, c(6, chains, nb_iter)), c(3, 2, 1))
  files <- character(chains)
  for (k in seq_len(chains)) {
    f <- tempfile(fileext = ".csv"); files[k] <- f
    lp <- d[, k, P]
    tab <- cbind(lp__ = lp, s[, k, ], d[, k, -P, drop = TRUE])
    colnames(tab) <- c("lp__", sp, gsub("\\[(\\d+)\\]", ".\\1", pn[-P]))
    con <- file(f, "w")
    writeLines(c("# model = fitoct_b200", paste0("# id = ", k), "# method = sample (Default)",
                 paste0("#     num_samples = ", nb_iter - nb_warmup), paste0("#     num_warmup = ", nb_warmup),
                 "#     save_warmup = 1", "#     thin = 1", "#     algorithm = hmc (Default)",
                 "#       engine = nuts (Default)"), con)
    write.table(tab, con, sep = ",", row.names = FALSE, quote = FALSE)
    writeLines(c("# Adaptation terminated", paste0("# Step size = ", res$stepsize[k]),
                 "# Diagonal elements of inverse mass matrix:",
                 paste0("# ", paste(matrix(res$inv_metric, ncol = chains)[, k], collapse = ", "))), con)
    close(con)
  }
  rstan::read_stan_csv(files)
}

fitExpGP <- function(x, y, uy, dataType = 2, Nn = 10, gridType = "internal", method = "sample",
                     theta0 = NULL, Sigma0 = NULL, lambda_rate = 0.1, rho_scale = 0.1,
                     nb_warmup = 500, nb_iter = 1500, prior_PD = 0, open_progress = FALSE,
                     chains = 4, seed = sample.int(.Machine$integer.max, 1)) {
  stopifnot(method %in% c("sample", "optim", "vb"))
  ctl <- list(dataType = dataType, Nn = Nn, gridType = as.integer(gridType == "extremal"),
              rho = ifelse(rho_scale == 0, 1 / Nn, rho_scale), lambda_rate = lambda_rate,
              theta0 = as.numeric(theta0), Sigma0 = as.numeric(Sigma0), prior_PD = prior_PD,
              chains = chains, nb_warmup = nb_warmup, nb_iter = nb_iter, seed = seed)
  dx  <- 1 / (Nn + 1)
  xGP <- if (gridType == "internal") seq(dx / 2, 1 - dx / 2, length.out = Nn) else seq(0, 1, length.out = Nn)
  if (method == "optim") {   # MAP + Hessian (MODEL_SPEC 10); fit$par$... as plotExpGP.R:13-17 reads it
    r <- .Call("foct_R_expgp_map", as.numeric(x), as.numeric(y), as.numeric(uy), ctl, PACKAGE = "FitOCTb200")
    p <- r$par
    fit <- list(par = list(theta = p[1:3], yGP = p[3 + seq_len(Nn)], lambda = p[Nn + 4], sigma = p[Nn + 5], br = p[Nn + 6],
                           m = r$m, resid = r$resid, dL = r$dL),
                value = p[Nn + 7], hessian = r$hessian, return_code = r$status)
    return(list(fit = fit, method = method, xGP = xGP, prior_PD = prior_PD))
  }
  if (method == "vb") {      # Stan's mean-field ADVI (MODEL_SPEC 14); the draws go through read_stan_csv like a 1-chain fit
    r <- .Call("foct_R_vb", as.numeric(x), as.numeric(y), as.numeric(uy), ctl, PACKAGE = "FitOCTb200")
    if (r$status == 2L) stop("fitExpGP(method='vb'): ADVI failed (dropped evaluations / no usable step size)")
    ns  <- length(r$draws) / (Nn + 7)
    res <- list(draws = r$draws, sampler_params = numeric(6 * ns), stepsize = r$eta, inv_metric = exp(2 * r$omega))
    fit <- .as_stanfit(res, 0L, Nn, 1L, 0L, ns)
    attr(fit, "vb") <- r[c("mean", "mu", "omega", "elbo", "eta", "iters", "status")]
    return(list(fit = fit, method = method, xGP = xGP, prior_PD = prior_PD))
  }
  res <- .Call("foct_R_sample", 0L, as.numeric(x), as.numeric(y), as.numeric(uy), ctl, PACKAGE = "FitOCTb200")
  # rstan-style progress lines so that the Shiny log scraper (ShinyInterface/server.R:457-472) reaches 100 %:
  # the whole fit is one kernel launch (~1.5 s), so only the final state of each chain is reported
  for (k in seq_len(chains))
    cat(sprintf("Chain %d: Iteration: %d / %d [100%%]  (Sampling)\n", k, nb_iter, nb_iter))
  list(fit = .as_stanfit(res, 0L, Nn, chains, nb_warmup, nb_iter), method = method, xGP = xGP, prior_PD = prior_PD)
}

fitMonoExp <- function(x, y, uy, dataType = 2) {
  r   <- .Call("foct_R_monoexp_map", as.numeric(x), as.numeric(y), as.numeric(uy), as.integer(dataType),
               PACKAGE = "FitOCTb200")
  cov <- solve(-r$hessian)
  list(best.theta = r$theta, cor.theta = cov2cor(cov),
       fit = list(par = list(theta = r$theta, m = r$m, resid = r$resid, br = r$br), hessian = r$hessian,
                  return_code = r$status),
       method = "optim")
}

# FitOCTLib::estimateNoise (FitOCT.R:89-91): smooth.spline at `df`, then uy = a_1 exp(-x/a_2) fitted to the residuals
estimateNoise <- function(x, y, df = 15, maxRate = 10000) {
  r <- .Call("foct_R_estimate_noise", as.numeric(x), as.numeric(y), as.numeric(df), as.numeric(maxRate),
             PACKAGE = "FitOCTb200")
  list(fit = list(par = list(theta = r$theta), spar = r$info[1], lambda = r$info[2], df = r$info[3], return_code = r$status),
       theta = r$theta, uy = r$uy, ySmooth = r$ySmooth, method = "optim")
}

# FitOCTLib::printBr (plotMonoExp.R:10, plotExpGP.R:22): NULL alert <=> fit OK (the gate at FitOCT.R:100)
printBr <- function(fit, silent = FALSE) {
  if (inherits(fit, "stanfit")) {
    br  <- mean(rstan::extract(fit, "br")[[1]])
    N   <- length(grep("^resid\\[", names(fit)))   # not saved by this back-end: pass the data length via attr(fit, "N")
    if (N == 0) N <- attr(fit, "N")
    np  <- length(grep("^(theta|yGP)\\[", names(fit)))
  } else {
    br <- fit$par$br; N <- length(fit$par$resid); np <- length(fit$par$theta) + length(fit$par$yGP)
  }
  ci    <- .Call("foct_R_birge_ci", as.numeric(N - np), PACKAGE = "FitOCTb200")
  alert <- if (br < ci[1] || br > ci[2]) "!!! WARNING !!! br out of interval" else NULL
  if (!silent) {
    cat("br   :", signif(br, 2), "\n"); cat("CI95 :", paste0(signif(ci, 2), collapse = "-"), "\n")
    if (!is.null(alert)) cat(alert, "\n")
  }
  list(br = br, CI95 = ci, alert = alert)
}

# FitOCTLib::estimateExpPrior (FitOCT.R:103-107)
estimateExpPrior <- function(x, uy, dataType, priorType = "mono", out, ru_theta = 0.05, eps = 1e-3) {
  fit <- out$fit
  .Call("foct_R_exp_prior", as.numeric(x), as.numeric(fit$par$m + fit$par$resid), as.numeric(uy), as.integer(dataType),
        as.integer(priorType == "abc"), as.numeric(out$best.theta), as.numeric(fit$hessian), as.numeric(ru_theta),
        PACKAGE = "FitOCTb200")
}
