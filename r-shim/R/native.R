# Drop-in replacements for the numeric bodies of getK() and BGGE() (reference: R/getK.R, R/BGGE.R).
# Signatures, validation, names and the returned objects are the reference's; the arithmetic is one
# .Call each.  NOT RUN in this repository's CI (no R in the build image).

getK <- function(Y, X, kernel = c("GK", "GB"), setKernel = NULL, bandwidth = 1,
                 model = c("SM", "MM", "MDs", "MDe"), quantil = 0.5, intercept.random = FALSE) {
  Y <- as.data.frame(Y)
  Y[colnames(Y)[1:2]] <- lapply(Y[colnames(Y)[1:2]], factor)
  subjects <- levels(Y[, 2]); env <- levels(Y[, 1]); nEnv <- length(env)
  if (any(table(Y[, c(1:2)]) > 1))
    warning("There are repeated genotypes in some environment. They were kept")
  if (model == "SM" && nEnv > 1)
    stop("Single model choosen, but more than one environment is in the phenotype file")
  if (model %in% c("MM", "MDs", "MDe") && nEnv < 2)
    stop("contrasts can be applied only to factors with 2 or more levels")
  if (!model %in% c("SM", "MM", "MDs", "MDe")) stop("Model selected is not available ")
  stopifnot(is.null(setKernel))                 # user kernels: bgge_getk_expand per kernel (omitted here)
  if (is.null(rownames(X))) stop("Genotype names are missing")
  if (!all(subjects %in% rownames(X)))
    stop("Not all genotypes presents in the phenotypic file are in marker matrix")
  X <- X[subjects, , drop = FALSE]; storage.mode(X) <- "double"
  kind <- switch(kernel, GB = 0L, GK = 1L,
                 stop("kernel selected is not available. Please choose one method available or make available other kernel through argument K"))
  mdl <- switch(model, SM = 0L, MM = 0L, MDs = 1L, MDe = 2L)
  mats <- .Call("bgge_R_getk", X, kind, as.double(bandwidth), as.double(quantil),
                as.integer(Y[, 2]) - 1L, as.integer(Y[, 1]) - 1L, nEnv, mdl, intercept.random)
  nb <- if (kind == 0L) 1L else length(bandwidth)
  tmp <- as.character(seq_len(nb))
  nm <- if (nb > 1) paste("G", tmp, sep = "_") else "G"
  ty <- rep("D", nb)
  if (mdl == 1L) { nm <- c(nm, if (nb > 1) paste("GE", tmp, sep = "_") else "GE"); ty <- c(ty, rep("BD", nb)) }
  if (mdl == 2L) { nm <- c(nm, if (nb > 1) paste(rep(env, nb), rep(tmp, each = nEnv), sep = "_") else env)
                   ty <- c(ty, rep("BD", nb * nEnv)) }
  if (intercept.random) { nm <- c(nm, "Gi"); ty <- c(ty, "D") }
  out <- Map(function(m, t) list(Kernel = m, Type = t), mats, ty)
  names(out) <- nm
  out
}

BGGE <- function(y, K, XF = NULL, ne, ite = 1000, burn = 200, thin = 3, verbose = FALSE, tol = 1e-10, R2 = 0.5) {
  if (is.null(names(K))) names(K) <- paste("K", seq(length(K)), sep = "")
  typeM <- sapply(K, function(x) x$Type)
  if (!all(typeM %in% c("BD", "D"))) stop("Matrix should be of types BD or D")
  if (length(ne) == 1 & any(typeM == "BD")) stop("Type BD should be used only for block diagonal matrices")
  r <- .Call("bgge_R_fit", as.numeric(y), lapply(K, function(x) x$Kernel), as.integer(typeM == "BD"),
             XF, as.double(ne), ite, burn, thin, tol, R2, sample.int(.Machine$integer.max, 1))
  out <- list(yHat = if (is.null(XF)) r$yHat else matrix(r$yHat, ncol = 1), varE = r$varE, varE.sd = r[["varE.sd"]])
  out$K <- setNames(lapply(seq_along(K), function(j)
    list(u = r$u[, j], u.sd = r[["u.sd"]][, j], varu = r$varu[j], varu.sd = r[["varu.sd"]][j])), names(K))
  out$chain <- list(mu = r[["chain.mu"]], varE = r[["chain.varE"]],
                    K = setNames(lapply(seq_along(K), function(j) list(varU = r[["chain.varU"]][, j])), names(K)))
  out$ite <- ite; out$burn <- burn; out$thin <- thin
  out$y <- r$y
  class(out) <- "BGGE"
  out
}
