# Drop-in replacements for the numeric bodies of getK() and BGGE() (reference: R/getK.R, R/BGGE.R).
# Signatures, validation, names and the returned objects are the reference's; the arithmetic is in libbgge_b200.
# NOT RUN in this repository's CI (no R in the build image); src/bgge_shim.c is type-checked against a mock R API.
#
# getK() returns the reference's list(Kernel = <n x n matrix>, Type = "D" | "BD") entries.  Every entry built from a
# base kernel also carries attr(entry, "bgge_factor") = the L x L base kernel with attributes mode / env_k / L / gid /
# env: BGGE() hands THAT to the library (setDEC through the L x L problem, factorised sweep) after checking, entry by
# entry, that `Kernel` still equals the expansion -- an edited matrix goes the dense route, like any hand-built K.
# options(BGGE.factored = TRUE): getK() skips the n x n matrices altogether (Kernel = NULL), for problems where they
# do not fit the host (config C5: 2 x 12.8 GB).
# options(BGGE.seed = k) fixes the device RNG seed (default: drawn from R's RNG, so set.seed() makes runs repeatable);
# options(BGGE.tape = list(normals = , chisq = )) injects recorded draws (SURVEY.md A.3, tests/golden/make_golden.R).

.bgge_entry <- function(K0, L, gid, env, mode, env_k, type, dense) {
  f <- if (is.null(K0)) numeric(0) else K0
  attr(f, "mode") <- as.integer(mode); attr(f, "env_k") <- as.integer(env_k); attr(f, "L") <- as.integer(L)
  attr(f, "gid") <- gid; attr(f, "env") <- env
  e <- list(Kernel = if (dense) .Call("bgge_R_expand", K0, as.integer(L), gid, env, as.integer(mode), as.integer(env_k)) else NULL,
            Type = type)
  attr(e, "bgge_factor") <- f
  e
}

getK <- function(Y, X, kernel = c("GK", "GB"), setKernel = NULL, bandwidth = 1,
                 model = c("SM", "MM", "MDs", "MDe"), quantil = 0.5, intercept.random = FALSE) {
  Y <- as.data.frame(Y)                                                                    # R/getK.R:99-105
  Y[colnames(Y)[1:2]] <- lapply(Y[colnames(Y)[1:2]], factor)
  subjects <- levels(Y[, 2]); env <- levels(Y[, 1]); nEnv <- length(env)
  if (any(table(Y[, c(1:2)]) > 1))                                                        # :108-109
    warning("There are repeated genotypes in some environment. They were kept")
  switch(model,                                                                           # :111-122, :188-229
         SM = if (nEnv > 1) stop("Single model choosen, but more than one environment is in the phenotype file"),
         MM = , MDs = , MDe = if (nEnv < 2) stop("contrasts can be applied only to factors with 2 or more levels"),
         stop("Model selected is not available "))
  gid <- as.integer(Y[, 2]) - 1L; ecd <- as.integer(Y[, 1]) - 1L; L <- length(subjects)
  if (is.null(setKernel)) {
    if (is.null(rownames(X))) stop("Genotype names are missing")                          # :125-126
    if (!all(subjects %in% rownames(X)))                                                  # :128-129
      stop("Not all genotypes presents in the phenotypic file are in marker matrix")
    X <- X[subjects, , drop = FALSE]; storage.mode(X) <- "double"                          # :131
    kind <- switch(kernel, GB = 0L, GK = 1L,                                               # :133-153
                   stop("kernel selected is not available. Please choose one method available or make available other kernel through argument K"))
    base <- .Call("bgge_R_getk_base", X, kind, as.double(bandwidth), as.double(quantil))
    tmp.names <- as.character(seq_along(base))                                            # :155
  } else {                                                                                # :157-182
    nullNames <- sapply(setKernel, function(x) any(sapply(dimnames(x), is.null)))
    if (any(nullNames)) stop("Genotype names are missing in some kernel")
    eqNames <- sapply(setKernel, function(x) all(subjects %in% rownames(x)) & all(subjects %in% colnames(x)))
    if (!all(eqNames)) stop("Not all genotypes presents in phenotypic file are in the kernel matrix.\n             Please check dimnames")
    base <- lapply(setKernel, function(x) { m <- x[subjects, subjects]; storage.mode(m) <- "double"; m })
    tmp.names <- if (is.null(names(base))) as.character(seq_along(base)) else names(base)
  }
  dense <- !isTRUE(getOption("BGGE.factored", FALSE))
  nb <- length(base)
  out <- lapply(base, function(K0) .bgge_entry(K0, L, gid, ecd, 0L, 0L, "D", dense))
  names(out) <- if (nb > 1) paste("G", tmp.names, sep = "_") else "G"                       # :184-185
  if (model == "MDs") {                                                                   # :198-204
    ge <- lapply(base, function(K0) .bgge_entry(K0, L, gid, ecd, 1L, 0L, "BD", dense))
    names(ge) <- if (nb > 1) paste("GE", tmp.names, sep = "_") else "GE"
    out <- c(out, ge)
  } else if (model == "MDe") {                                                            # :206-226
    for (b in seq_len(nb)) {
      ek <- lapply(seq_len(nEnv), function(k) .bgge_entry(base[[b]], L, gid, ecd, 2L, k - 1L, "BD", dense))
      names(ek) <- if (nb > 1) paste(env, tmp.names[b], sep = "_") else env
      out <- c(out, ek)
    }
  }
  if (intercept.random) out <- c(out, list(Gi = .bgge_entry(NULL, L, gid, ecd, 3L, 0L, "D", dense)))   # :231-234
  out
}

BGGE <- function(y, K, XF = NULL, ne, ite = 1000, burn = 200, thin = 3, verbose = FALSE, tol = 1e-10, R2 = 0.5) {
  if (is.null(names(K))) names(K) <- paste("K", seq(length(K)), sep = "")                 # R/BGGE.R:203-205
  typeM <- sapply(K, function(x) x$Type)
  if (!all(typeM %in% c("BD", "D"))) stop("Matrix should be of types BD or D")            # :219-220
  if (length(ne) == 1 & any(typeM == "BD")) stop("Type BD should be used only for block diagonal matrices")   # :222-223
  gid <- integer(0); env <- integer(0)
  kernels <- lapply(K, function(k) {
    f <- attr(k, "bgge_factor")
    if (!is.null(f)) {
      ok <- is.null(k$Kernel) ||
        .Call("bgge_R_factor_matches", k$Kernel, if (length(f)) f else NULL, attr(f, "L"), attr(f, "gid"), attr(f, "env"),
              attr(f, "mode"), attr(f, "env_k"))
      if (ok) { gid <<- attr(f, "gid"); env <<- attr(f, "env"); return(f) }
    }
    m <- k$Kernel; storage.mode(m) <- "double"; m                                          # dense route (:144, :160)
  })
  tape <- getOption("BGGE.tape", NULL)
  seed <- getOption("BGGE.seed", NULL)
  if (is.null(seed)) seed <- sample.int(.Machine$integer.max, 1)
  if (!is.null(XF)) { XF <- as.matrix(XF); storage.mode(XF) <- "double" }
  r <- .Call("bgge_R_fit", as.numeric(y), kernels, as.integer(typeM == "BD"), gid, env, XF, as.double(ne),
             ite, burn, thin, as.integer(verbose), tol, R2, as.double(seed),
             if (is.null(tape)) NULL else as.double(tape$normals), if (is.null(tape)) NULL else as.double(tape$chisq))
  out <- list(yHat = if (is.null(XF)) r$yHat else matrix(r$yHat, ncol = 1), varE = r$varE, varE.sd = r[["varE.sd"]])   # :410-443
  out$K <- setNames(lapply(seq_along(K), function(j)
    list(u = r$u[, j], u.sd = r[["u.sd"]][, j], varu = r$varu[j], varu.sd = r[["varu.sd"]][j])), names(K))
  out$chain <- list(mu = r[["chain.mu"]], varE = r[["chain.varE"]],
                    K = setNames(lapply(seq_along(K), function(j) list(varU = r[["chain.varU"]][, j])), names(K)))
  out$ite <- ite; out$burn <- burn; out$thin <- thin
  out$y <- r$y
  class(out) <- "BGGE"
  out
}
