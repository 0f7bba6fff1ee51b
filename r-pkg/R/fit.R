# Drop-in wrappers with the signatures the reference calls (FitOCT.R:95, 110-124; priPost.R:2-16;
# ShinyInterface/server.R:341-343, 408-426).  Swap `FitOCTLib::fitExpGP` for `FitOCTb200::fitExpGP`.
# NOT RUN in this repository's image (no R): kept logic-free; all numerics are behind .Call.

.par_names <- function(kind, Nn) {
  if (kind == 0L) c(paste0("theta[", 1:3, "]"), paste0("yGP[", seq_len(Nn), "]"), "lambda", "sigma", "br", "lp__")
  else c(paste0("theta[", 1:3, "]"), "br", "lp__")
}

# Build a genuine rstan stanfit through rstan's own reader: one Stan-CSV per chain -> rstan::read_stan_csv.
# gq: optional list(m, resid, dL), each [N, nb_iter * chains] in the row order of the draws (iteration-major, chain
# fastest), appended as the columns m[i], resid[i], dL[i] rstan would have saved (plotExpGP.R:55-58 hands the fit to
# FitOCTLib::plotExpGP, which reads them).
.as_stanfit <- function(res, kind, Nn, chains, nb_warmup, nb_iter, gq = NULL) {
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
    if (!is.null(gq)) {
      rows <- (seq_len(nb_iter) - 1) * chains + k
      for (nm in c("m", "resid", "dL")) {
        g <- t(gq[[nm]][, rows, drop = FALSE])
        colnames(g) <- paste0(nm, ".", seq_len(ncol(g)))
        tab <- cbind(tab, g)
      }
    }
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

.ctl <- function(dataType, Nn, gridType, rho_scale, lambda_rate, theta0, Sigma0, prior_PD, chains, nb_warmup, nb_iter,
                 seed, rhat_target = 0, max_extend = 0, ...) {
  c(list(dataType = dataType, Nn = Nn, gridType = as.integer(gridType == "extremal"),
         rho = ifelse(rho_scale == 0, 1 / Nn, rho_scale), lambda_rate = lambda_rate,
         theta0 = as.numeric(theta0), Sigma0 = as.numeric(Sigma0), prior_PD = prior_PD,
         chains = chains, nb_warmup = nb_warmup, nb_iter = nb_iter, seed = seed,
         rhat_target = rhat_target, max_extend = max_extend), list(...))
}

# gq = TRUE adds the generated quantities m[i], resid[i], dL[i] of every saved draw to the stanfit (what the Stan model's
# generated quantities block gives FitOCTLib::plotExpGP / plotPriPostAll, plotExpGP.R:55-58, priPost.R:22); off by
# default because it is 3 N columns per draw.  rhat_target / max_extend: continue the chains on the device until the
# largest split R-hat is below the target (include/fitoct_b200.h).  Progress lines ("Chain k: Iteration: ...") are
# printed by the shim while the kernels run, and Ctrl-C / the Shiny stop button cancels them.
fitExpGP <- function(x, y, uy, dataType = 2, Nn = 10, gridType = "internal", method = "sample",
                     theta0 = NULL, Sigma0 = NULL, lambda_rate = 0.1, rho_scale = 0.1,
                     nb_warmup = 500, nb_iter = 1500, prior_PD = 0, open_progress = FALSE,
                     chains = 4, seed = sample.int(.Machine$integer.max, 1), gq = FALSE,
                     rhat_target = 0, max_extend = 0) {
  stopifnot(method %in% c("sample", "optim", "vb"))
  stopifnot(!is.null(theta0), length(theta0) == 3, !is.null(Sigma0), length(Sigma0) == 9)
  x <- as.numeric(x); y <- as.numeric(y); uy <- as.numeric(uy)
  ctl <- .ctl(dataType, Nn, gridType, rho_scale, lambda_rate, theta0, Sigma0, prior_PD, chains, nb_warmup, nb_iter, seed,
              rhat_target, max_extend)
  dx  <- 1 / (Nn + 1)
  xGP <- if (gridType == "internal") seq(dx / 2, 1 - dx / 2, length.out = Nn) else seq(0, 1, length.out = Nn)
  if (method == "optim") {   # MAP + Hessian (MODEL_SPEC 10); fit$par$... as plotExpGP.R:13-17 reads it
    r <- .Call("foct_R_expgp_map", x, y, uy, ctl, PACKAGE = "FitOCTb200")
    p <- r$par
    fit <- list(par = list(theta = p[1:3], yGP = p[3 + seq_len(Nn)], lambda = p[Nn + 4], sigma = p[Nn + 5], br = p[Nn + 6],
                           m = r$m, resid = r$resid, dL = r$dL),
                value = p[Nn + 7], hessian = r$hessian, return_code = r$status)
    return(list(fit = fit, method = method, xGP = xGP, prior_PD = prior_PD))
  }
  if (method == "vb") {      # Stan's mean-field ADVI (MODEL_SPEC 14); the draws go through read_stan_csv like a 1-chain fit
    r <- .Call("foct_R_vb", x, y, uy, ctl, PACKAGE = "FitOCTb200")
    if (r$status == 2L) stop("fitExpGP(method='vb'): ADVI failed (dropped evaluations / no usable step size)")
    ns  <- length(r$draws) / (Nn + 7)
    res <- list(draws = r$draws, sampler_params = numeric(6 * ns), stepsize = r$eta, inv_metric = exp(2 * r$omega))
    g   <- if (gq) .Call("foct_R_predict", 0L, x, y, uy, ctl, r$draws, PACKAGE = "FitOCTb200") else NULL
    fit <- .as_stanfit(res, 0L, Nn, 1L, 0L, ns, g)
    attr(fit, "vb") <- r[c("mean", "mu", "omega", "elbo", "eta", "iters", "status")]
    attr(fit, "N") <- length(x)   # printBr needs the data length when resid[] is not in the fit (plotExpGP.R:22)
    return(list(fit = fit, method = method, xGP = xGP, prior_PD = prior_PD))
  }
  res <- .Call("foct_R_sample", 0L, x, y, uy, ctl, PACKAGE = "FitOCTb200")
  g   <- if (gq) .Call("foct_R_predict", 0L, x, y, uy, ctl, res$draws, PACKAGE = "FitOCTb200") else NULL
  fit <- .as_stanfit(res, 0L, Nn, chains, nb_warmup, nb_iter, g)
  attr(fit, "N") <- length(x)
  attr(fit, "n_extend") <- res$n_extend
  list(fit = fit, method = method, xGP = xGP, prior_PD = prior_PD)
}

# Batch form (SURVEY 8b / 8f-N4): every profile of a directory in ONE .Call, sharded over `devices` (CUDA device numbers,
# NULL = current device).  xs, ys, uys: lists of numeric vectors; theta0: 3 x n matrix; Sigma0: 9 x n matrix (each
# column a 3 x 3 covariance).  Returns one fitExpGP-style list per profile; stanfit = FALSE keeps the raw arrays
# (draws [P_out, chains, n_saved], summary [11, P_out]) for 1e5-profile batches where 1e5 stanfit objects are not wanted.
fitExpGP_batch <- function(xs, ys, uys, theta0, Sigma0, dataType = 2, Nn = 10, gridType = "internal", lambda_rate = 0.1,
                           rho_scale = 0.1, nb_warmup = 500, nb_iter = 1500, prior_PD = 0, chains = 4,
                           seed = sample.int(.Machine$integer.max, 1), devices = NULL, stanfit = TRUE, draws = TRUE,
                           rhat_target = 0, max_extend = 0, quiet = FALSE) {
  n <- length(xs)
  stopifnot(length(ys) == n, length(uys) == n, length(theta0) == 3 * n, length(Sigma0) == 9 * n)
  ctl <- .ctl(dataType, Nn, gridType, rho_scale, lambda_rate, theta0, Sigma0, prior_PD, chains, nb_warmup, nb_iter, seed,
              rhat_target, max_extend, draws = as.integer(draws), quiet = as.integer(quiet))
  r <- .Call("foct_R_sample_batch", 0L, lapply(xs, as.numeric), lapply(ys, as.numeric), lapply(uys, as.numeric), ctl,
             if (is.null(devices)) NULL else as.integer(devices), PACKAGE = "FitOCTb200")
  P   <- Nn + 7; D <- Nn + 5
  dx  <- 1 / (Nn + 1)
  xGP <- if (gridType == "internal") seq(dx / 2, 1 - dx / 2, length.out = Nn) else seq(0, 1, length.out = Nn)
  summ <- array(r$summary, c(11, P, n))
  dimnames(summ) <- list(c("mean", "se_mean", "sd", "2.5%", "25%", "50%", "75%", "97.5%", "n_eff", "Rhat", "Bulk_ESS"),
                         .par_names(0L, Nn), NULL)
  dr <- if (draws) array(r$draws, c(P, chains, nb_iter, n)) else NULL
  sp <- if (draws) array(r$sampler_params, c(6, chains, nb_iter, n)) else NULL
  lapply(seq_len(n), function(j) {
    one <- list(draws = if (draws) dr[, , , j] else NULL, sampler_params = if (draws) sp[, , , j] else NULL,
                stepsize = matrix(r$stepsize, chains)[, j], inv_metric = array(r$inv_metric, c(D, chains, n))[, , j])
    fit <- if (stanfit && draws) .as_stanfit(one, 0L, Nn, chains, nb_warmup, nb_iter) else one
    attr(fit, "N") <- length(xs[[j]])
    list(fit = fit, method = "sample", xGP = xGP, prior_PD = prior_PD, summary = t(summ[, , j]), n_extend = r$n_extend[j])
  })
}

# The body of FitOCT.R's dataset loop (FitOCT.R:84-124) for ALL datasets in one call: estimateNoise -> fitMonoExp ->
# printBr gate -> estimateExpPrior -> fitExpGP on the profiles the gate lets through.  xs, ys: what selX returned for
# each Courbe.csv.  The arguments are ctrlParams.yaml's keys (FitOCT.R:37-53).
FitOCT_batch <- function(xs, ys, dataType = 2, smooth_df = 15, priorType = "abc", ru_theta = 0.05, Nn = 10,
                         gridType = "internal", rho_scale = 0, lambda_rate = 0.1, nb_warmup = 500, nb_sample = 1000,
                         chains = 4, seed = 1234, devices = NULL, gate = TRUE, draws = FALSE, rhat_target = 0, max_extend = 0) {
  n <- length(xs)
  ctl <- list(dataType = dataType, smooth_df = smooth_df, priorType = as.integer(priorType == "abc"), ru_theta = ru_theta,
              Nn = Nn, gridType = as.integer(gridType == "extremal"), rho_scale = rho_scale, lambda_rate = lambda_rate,
              nb_warmup = nb_warmup, nb_iter = nb_warmup + nb_sample, chains = chains, seed = seed, gate = as.integer(gate),
              draws = as.integer(draws), rhat_target = rhat_target, max_extend = max_extend)
  r  <- .Call("foct_R_pipeline", lapply(xs, as.numeric), lapply(ys, as.numeric), ctl,
              if (is.null(devices)) NULL else as.integer(devices), PACKAGE = "FitOCTb200")
  Ns <- vapply(xs, length, 1L); off <- c(0L, cumsum(Ns))
  P  <- Nn + 7
  k  <- r$n_expgp
  summ <- array(r$summary, c(11, P, n))[, , seq_len(k), drop = FALSE]
  lapply(seq_len(n), function(j) {
    idx <- (off[j] + 1):off[j + 1]
    e   <- match(j - 1L, r$expgp_index[seq_len(k)])
    list(uy = r$uy[idx], ySmooth = r$ySmooth[idx], noise_theta = r$noise_theta[2 * j - 1:0],
         best.theta = r$mono_theta[3 * j - 2:0], hessian = matrix(r$mono_hessian[9 * j - 8:0], 3), br = r$mono_br[j],
         CI95 = r$br_ci[2 * j - 1:0], alert = if (r$alert[j]) "!!! WARNING !!! br out of interval" else NULL,
         theta0 = r$theta0[3 * j - 2:0], Sigma0 = matrix(r$Sigma0[9 * j - 8:0], 3),
         expgp_summary = if (is.na(e)) NULL else t(summ[, , e]))
  })
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

# rstan's summary(fit)$summary / monitor() for draws the caller holds (an [iterations, chains, parameters] array, the
# shape of as.array(stanfit)): mean, se_mean, sd, quantiles, n_eff, Rhat, Bulk_ESS on the device (foct_summary).
summaryDraws <- function(draws) {
  stopifnot(length(dim(draws)) == 3, dim(draws)[1] >= 4)
  d <- dim(draws)
  s <- .Call("foct_R_summary", as.numeric(aperm(draws, c(3, 2, 1))), as.integer(d[1]), as.integer(d[2]), PACKAGE = "FitOCTb200")
  dimnames(s) <- list(dimnames(draws)[[3]], c("mean", "se_mean", "sd", "2.5%", "25%", "50%", "75%", "97.5%", "n_eff", "Rhat", "Bulk_ESS"))
  s
}
