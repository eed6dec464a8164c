# epimethex.analysis — same nine arguments, same output directory "<min>-<max>", same file name and columns as
# the reference (R/EpimethexAnalysis.R:70-72, :88-91, :470; R/Functions.R:224-229). The string work stays in R
# (annotation explode, R's own collation for ordering); integers and doubles go to libepx through .Call.
# NOT RUN in the build environment (no R there); epimethex_b200/analysis.py is the executed mirror of this file.
# device: one GPU (0L) or several (0:7): R stays one process, libepx drives every GPU from its own host thread.
epimethex.analysis <- function(dfExpressions, dfGpl, dfMethylation, minRangeGene, maxRangeGene, numCores,
                               dataGenesLinear, testExpression, testMethylation,
                               session = NULL, compat_fc = c("reference", "means"), strict = FALSE, device = 0L) {
    compat_fc <- match.arg(compat_fc)
    # --- R/EpimethexAnalysis.R:86, :103-114: slice, drop all-zero genes, drop samples with NA ---
    dfExpressions <- dfExpressions[minRangeGene:maxRangeGene, ]
    nameFD <- paste(as.character(minRangeGene), as.character(maxRangeGene), sep = "-")
    dir.create(file.path(getwd(), nameFD), showWarnings = FALSE)
    genes <- if ("sample" %in% colnames(dfExpressions)) as.character(dfExpressions$sample) else rownames(dfExpressions)
    x <- t(as.matrix(dfExpressions[, -1]))                       # samples x genes
    keep <- which(!apply(x == 0, 2, all))
    x <- x[, keep, drop = FALSE]; genes <- genes[keep]
    x <- x[stats::complete.cases(x), , drop = FALSE]
    samples <- rownames(x)
    storage.mode(x) <- "double"                                  # the shim takes doubles only (integer counts arrive as INTSXP)
    # --- :93-100, :248-265, :287-291: annotation explode, keep analysed genes, order, dedup on (cg, position) ---
    dfGpl <- dfGpl[!(is.na(dfGpl$UCSC_RefGene_Name) | dfGpl$UCSC_RefGene_Name == ""), ]
    island <- paste(dfGpl$Relation_to_UCSC_CpG_Island, dfGpl$UCSC_CpG_Islands_Name, sep = "_")
    s <- strsplit(dfGpl$UCSC_RefGene_Name, ";"); s2 <- strsplit(dfGpl$UCSC_RefGene_Group, ";")
    k <- lengths(s)
    ann <- data.frame(cg = rep(dfGpl$ID, k), coord = rep(dfGpl$Coordinate_36, k), gene = unlist(s),
                      position = unlist(s2), island = rep(island, k), stringsAsFactors = FALSE)
    ann <- ann[ann$gene %in% genes, ]
    ann <- ann[order(ann$gene, ann$coord), ]                     # plyr::arrange: R's collation decides
    ann <- ann[!duplicated(ann[, c("cg", "position")]), ]        # dfCGunique (gene ignored on purpose, SURVEY 8g-6)
    # --- :278-282: methylation, incomplete probes dropped, columns matched to the expression samples by name ---
    dfMethylation <- stats::na.omit(dfMethylation)
    ids <- as.character(dfMethylation[[1]])
    if (is.null(session)) session <- .Call(C_epx_upload, unclass(dfMethylation[, samples, drop = FALSE]), as.integer(device))
    # --- CSR in expression order of the genes; output order = sorted gene names (reference: factor levels) ---
    gi <- match(ann$gene, genes)
    ord <- order(gi)                                             # stable: keeps dfCGunique order inside a gene
    rowPtr <- c(0, cumsum(tabulate(gi, nbins = length(genes))))
    probeIdx <- match(ann$cg[ord], ids) - 1L                     # NA = annotated probe without data
    res <- .Call(C_epx_analyze, session, x, as.numeric(rowPtr), as.integer(probeIdx),
                 as.integer(c(dataGenesLinear, testExpression, testMethylation, compat_fc == "means", strict)))
    back <- order(ord)                                           # reference row -> CSR pair
    tab <- rbind(res$pair[, back, drop = FALSE], res$gene[1:6, gi, drop = FALSE])
    rownames(tab) <- c("medianDown", "medianMedium", "medianUP", "bd_UPvsMID", "bd_UPvsDOWN", "bd_MIDvsDOWN",
                       "pvalue_UPvsMID", "pvalue_UPvsDOWN", "pvalue_MIDvsDOWN", "pearson_correlation",
                       "pvalue_pearson_correlation", "fc_UPvsMID(gene)", "fc_UPvsDOWN(gene)", "fc_MIDvsDOWN(gene)",
                       "pvalue_UPvsMID(gene)", "pvalue_UPvsDOWN(gene)", "pvalue_MIDvsDOWN(gene)")
    colnames(tab) <- paste(ann$cg, ann$position, ann$gene, sep = "_")
    out <- rbind(tab, island = ann$island)                       # the character row coerces all, as in :463-467
    utils::write.csv(t(out), file.path(getwd(), nameFD, "CG_data_individually.csv"))
    # --- pooled tables: Loops B, C, D (R/EpimethexAnalysis.R:476-620, R/Functions.R:136-188) ---
    csr <- back - 1L                                             # 0-based pair index of every dfCGunique row
    hasData <- !is.na(probeIdx[back])
    gfac <- factor(ann$gene)                                     # column order of the reference: sorted gene names
    pooled <- function(groups, names, file) {                    # groups: list of dfCGunique row numbers with data
        if (length(groups) == 0) { warning(paste(file, "is not generated")); return(invisible()) }
        ptr <- c(0, cumsum(lengths(groups)))
        g <- .Call(C_epx_groups, session, as.numeric(ptr), as.integer(csr[unlist(groups)]))
        first <- vapply(seq_along(groups), function(i) if (length(groups[[i]])) groups[[i]][1] else NA_integer_, 1L)
        tail6 <- res$gene[1:6, gi[first], drop = FALSE]
        tail6[, is.na(first)] <- NA                               # :536-541: a gene without data -> NA column
        t2 <- rbind(g, tail6); rownames(t2) <- rownames(tab); colnames(t2) <- names
        utils::write.csv(t(t2), file.path(getwd(), nameFD, file))
    }
    rows <- split(seq_len(nrow(ann)), gfac)
    pooled(lapply(rows, function(r) r[hasData[r]]), paste("CG", levels(gfac), sep = "~"), "CG_of_a_genes.csv")
    for (spec in list(list("position", unique(ann$position), "CG_by_position.csv"),
                      list("island", Filter(function(v) !grepl("NA_NA|^_", v), unique(ann$island)), "CG_by_islands.csv"))) {
        grp <- list(); nms <- character(0)
        for (gname in levels(gfac)) for (v in spec[[2]]) {       # gene-major, values in order of first appearance
            r <- rows[[gname]]; r <- r[ann[[spec[[1]]]][r] == v & hasData[r]]
            if (length(r) > 1) { grp[[length(grp) + 1]] <- r; nms <- c(nms, paste(v, gname, sep = "_")) }
        }
        pooled(grp, nms, spec[[3]])
    }
    invisible(session)
}
