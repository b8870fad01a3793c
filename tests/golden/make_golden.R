# make_golden.R -- records what the UNMODIFIED reference package (italo-granato/BGGE) computes, so that the oracle
# (oracle/bgge_oracle.py) and the CUDA path (bgge_b200) can be pinned to the reference itself.
#
# NOT RUN in this repository's build image (it has no R).  A maintainer with R >= 3.1.1 runs, from the repo root:
#
#     Rscript tests/golden/make_golden.R /path/to/BGGE        # source tree of the reference (uses R/getK.R, R/BGGE.R)
#     # or, with the package installed:  Rscript tests/golden/make_golden.R
#
# and commits the tests/golden/r_*.json it writes.  tests/test_golden_r.py then checks, for every file present:
#   * oracle.getK and bgge_b200.getK against the recorded kernels (<= 1e-12 relative),
#   * oracle.BGGE and bgge_b200.BGGE, fed the recorded draws and R's own eigen() output, against the recorded
#     chains and summaries (<= 1e-9 relative)  -- north_star's injected-draw test.
# Until such a file exists the parity of this repository stays "unpinned" (DESIGN.md section 2).
#
# How the draws are captured without editing the package (SURVEY.md A.3): BGGE() is copied and its enclosing
# environment is given shadow versions of rnorm / rgamma / eigen that call the real ones and record
#   normals : the standard normals z behind every rnorm(n, mean, sd) (the shadow returns mean + sd * z, which is
#             exactly what R's C code computes),
#   chisq   : for every rgamma(1, shape, rate) draw g the equivalent chi-square value 2 * rate * g
#             (1 / g == (2 * rate) / chisq, so sigma^2 = (SS + prior) / chisq_df as the tape contract states),
#   eigen   : every eigen() result in call order (values, vectors) -- the sampler's path depends on LAPACK's
#             eigenvector signs, so the comparison must run on R's own basis.

args <- commandArgs(trailingOnly = TRUE)
if (length(args) >= 1) {
  source(file.path(args[1], "R", "getK.R"))
  source(file.path(args[1], "R", "BGGE.R"))
} else {
  library(BGGE)
}
out_dir <- file.path("tests", "golden")
if (!dir.exists(out_dir)) stop("run from the repository root (tests/golden not found)")

# ---- minimal JSON writer (base R only; doubles with 17 significant digits, NA -> null) ------------------------------
num <- function(x) {
  s <- formatC(as.numeric(x), digits = 17, format = "g")
  s[is.na(x)] <- "null"
  s[is.infinite(x) & x > 0] <- "1e999"
  s[is.infinite(x) & x < 0] <- "-1e999"
  s
}
js <- function(x) {
  if (is.null(x)) return("null")
  if (is.list(x)) {
    inner <- vapply(x, js, character(1))
    if (!is.null(names(x)) && all(nzchar(names(x))))
      return(paste0("{", paste0("\"", names(x), "\": ", inner, collapse = ", "), "}"))
    return(paste0("[", paste(inner, collapse = ", "), "]"))
  }
  if (is.character(x)) return(if (length(x) == 1) paste0("\"", x, "\"") else paste0("[", paste0("\"", x, "\"", collapse = ", "), "]"))
  if (is.logical(x)) return(if (length(x) == 1) tolower(as.character(x)) else paste0("[", paste(tolower(as.character(x)), collapse = ", "), "]"))
  if (is.matrix(x))     # row-major list of rows, like numpy's tolist()
    return(paste0("[", paste(apply(x, 1, function(r) paste0("[", paste(num(r), collapse = ", "), "]")), collapse = ", "), "]"))
  if (length(x) == 1) return(num(x))
  paste0("[", paste(num(x), collapse = ", "), "]")
}

# ---- synthetic data of the reference's test shape, scaled down (tests/testthat/test.R:7-10, 37-42) ------------------
make_data <- function(L, p, E, seed, n_na) {
  set.seed(seed)
  maf <- runif(p, 0.05, 0.5)
  X <- sapply(maf, function(f) rbinom(L, 2, f))
  while (any(mono <- apply(X, 2, sd) == 0)) X[, mono] <- sapply(maf[mono], function(f) rbinom(L, 2, f))
  X <- scale(X)
  attr(X, "scaled:center") <- NULL; attr(X, "scaled:scale") <- NULL
  rownames(X) <- sprintf("%05d", 1:L)
  env <- rep(sprintf("env%02d", 1:E), each = L)
  gid <- rep(rownames(X), E)
  g <- as.numeric(X %*% rnorm(p)) / sqrt(p)
  y <- rep(rnorm(E, 0, 0.3), each = L) + rep(g, E) + rnorm(L * E, 0, 0.7)
  if (n_na > 0) y[sample(L * E, n_na)] <- NA
  list(X = X, Y = data.frame(env = env, gid = gid, y = y, stringsAsFactors = FALSE), y = y)
}

# ---- one recorded run -----------------------------------------------------------------------------------------------
record <- function(tag, L, p, E, kernel, model, ite, burn, thin, seed, n_na = 0, with_xf = FALSE, bandwidth = 1,
                   intercept.random = FALSE) {
  d <- make_data(L, p, E, seed, n_na)
  K <- getK(Y = d$Y, X = d$X, kernel = kernel, model = model, bandwidth = bandwidth, intercept.random = intercept.random)
  XF <- if (with_xf) model.matrix(~ factor(d$Y$env) - 1) else NULL
  rec <- new.env()
  rec$z <- numeric(0); rec$c <- numeric(0); rec$eig <- list()
  shadow <- list(
    rnorm = function(n, mean = 0, sd = 1) {
      z <- stats::rnorm(n)
      rec$z <- c(rec$z, z)
      mean + sd * z
    },
    rgamma = function(n, shape, rate = 1) {
      g <- stats::rgamma(n, shape, rate)
      rec$c <- c(rec$c, 2 * rate * g)
      g
    },
    eigen = function(x, ...) {
      r <- base::eigen(x, ...)
      rec$eig[[length(rec$eig) + 1]] <- list(values = r$values, vectors = r$vectors)
      r
    })
  f <- BGGE
  environment(f) <- list2env(shadow, parent = environment(BGGE))
  set.seed(seed + 1000)
  fit <- f(y = d$y, K = K, XF = XF, ne = rep(L, E), ite = ite, burn = burn, thin = thin)
  stopifnot(length(fit) == 9, inherits(fit, "BGGE"))
  obj <- list(
    tag = tag, reference = "italo-granato/BGGE", r_version = R.version.string,
    package_version = tryCatch(as.character(utils::packageVersion("BGGE")), error = function(e) "source"),
    L = L, p = p, E = E, kernel = kernel, model = model, bandwidth = bandwidth, intercept_random = intercept.random,
    ite = ite, burn = burn, thin = thin, ne = rep(L, E), tol = 1e-10, R2 = 0.5,
    X = unname(d$X), rownames = rownames(d$X), env = d$Y$env, gid = d$Y$gid, y = d$y,
    XF = if (is.null(XF)) NULL else unname(XF),
    K = lapply(K, function(k) list(Type = k$Type, Kernel = unname(k$Kernel))),
    kernel_names = names(K),
    tape = list(normals = rec$z, chisq = rec$c),
    eigen = rec$eig,
    fit = list(yHat = as.numeric(fit$yHat), varE = fit$varE, varE_sd = fit$varE.sd,
               K = lapply(fit$K, function(k) list(u = as.numeric(k$u), u_sd = as.numeric(k$u.sd), varu = k$varu, varu_sd = k$varu.sd)),
               chain = list(mu = fit$chain$mu, varE = fit$chain$varE, K = lapply(fit$chain$K, function(k) list(varU = k$varU))),
               y = as.numeric(fit$y)))
  path <- file.path(out_dir, paste0("r_", tag, ".json"))
  writeLines(js(obj), path)
  cat("wrote", path, " (", length(rec$z), "normals,", length(rec$c), "chi-squares,", length(rec$eig), "eigen calls )\n")
}

record("gb_mm",   L = 30, p = 50, E = 3, kernel = "GB", model = "MM",  ite = 30, burn = 6, thin = 2, seed = 1)
record("gk_mds",  L = 30, p = 50, E = 3, kernel = "GK", model = "MDs", ite = 30, burn = 6, thin = 2, seed = 2, n_na = 5)
record("gb_mde",  L = 24, p = 40, E = 4, kernel = "GB", model = "MDe", ite = 24, burn = 4, thin = 3, seed = 3, n_na = 3, with_xf = TRUE)
record("gk_mds_bw", L = 24, p = 40, E = 2, kernel = "GK", model = "MDs", ite = 20, burn = 2, thin = 1, seed = 4,
       bandwidth = c(0.5, 2), intercept.random = TRUE)
