# baseline/run_rstan.R — the TRUE reference number: real FitOCTLib::fitExpGP (rstan NUTS) on the same CSV
# inputs bench.py uses.  Cannot be run in this repository's image (no R / rstan / FitOCTLib, SURVEY.md F4);
# shipped so that someone with R can fill in the "rstan on host cores" column of BASELINE.md §4 and diff the
# posterior summaries against fitoct_b200's (3 MCSE, R-hat < 1.01).
#
#   python -c "from fitoct_b200 import synth; import numpy as np; S=synth.make_profiles(16, modulated_only=True); \
#              [np.savetxt(f'baseline/in/Courbe_{j}.csv', np.c_[S['x'],S['Y'][j],S['UY'][j]], delimiter=',', \
#               header='x,y,uy', comments='') for j in range(16)]; np.savetxt('baseline/in/theta0.csv', S['theta0'], delimiter=',')"
#   Rscript baseline/run_rstan.R baseline/in baseline/out
library(parallel); library(rstan); library(FitOCTLib)
options(mc.cores = parallel::detectCores()); rstan_options(auto_write = TRUE); set.seed(1234)   # FitOCT.R:13-15
args <- commandArgs(trailingOnly = TRUE); indir <- args[1]; outdir <- args[2]; dir.create(outdir, showWarnings = FALSE)
theta0 <- as.matrix(read.csv(file.path(indir, "theta0.csv"), header = FALSE))
files  <- sort(list.files(indir, pattern = "^Courbe_", full.names = TRUE))
Nn <- 10; nb_warmup <- 500; nb_sample <- 1000                                                     # FitOCT.R:43-49
t0 <- Sys.time()
for (j in seq_along(files)) {
  D <- read.csv(files[j])
  th <- theta0[j, ]
  fit <- FitOCTLib::fitExpGP(D$x, D$y, D$uy, dataType = 2, Nn = Nn, gridType = "internal", method = "sample",
                             theta0 = th, Sigma0 = diag((0.05 * th)^2), lambda_rate = 0.1, rho_scale = 1 / Nn,
                             nb_warmup = nb_warmup, nb_iter = nb_warmup + nb_sample, prior_PD = 0,
                             open_progress = FALSE)                                               # FitOCT.R:110-124
  s <- rstan::summary(fit$fit, pars = c("theta", "yGP", "lambda", "sigma", "br", "lp__"))$summary
  write.csv(s, file.path(outdir, sprintf("summary_%d.csv", j - 1)))
  write.csv(as.matrix(fit$fit), file.path(outdir, sprintf("draws_%d.csv", j - 1)), row.names = FALSE)
}
wall <- as.numeric(difftime(Sys.time(), t0, units = "secs"))
cat(sprintf("profiles=%d cores=%d wall_s=%.1f draws_per_s=%.1f\n", length(files), detectCores(), wall,
            length(files) * 4 * nb_sample / wall))
