# r/ECS_b200.R -- the R side of the drop-in: replacement definitions for the four set-level wrappers of ClustAssess's
# R/ECS.R that loop over partition pairs in R today.  Every exported signature, argument check, return shape and `names`
# behaviour is the reference's; only the pair loops are gone: each wrapper makes ONE .Call into the B200 library through
# the Rcpp routines of r/calculate_ecs_b200.cpp (route A of INTEGRATION.md; route B binds the same entries under the
# _ClustAssess_ecs_* names of r/clustassess_shim.c).
#
#   replaces (reference file:line)                               calls
#   element_sim_matrix_flat_disjoint   R/ECS.R:931-981           ecsSimMatrix(mb_list)
#   merge_identical_partitions         R/ECS.R:1021-1054         ecsMergeIdentical(mb_list, freq)
#   weighted_element_consistency       R/ECS.R:1446-1500         ecsConsistency(mb_list, weights, calculate_sim_matrix)
#   element_agreement_flat_disjoint    R/ECS.R:1625-1648         ecsAgreement(reference, mb_list)
#
# r/apply_ecs_b200.py splices these definitions into a ClustAssess checkout (it replaces exactly those four function
# bodies and leaves every other line of R/ECS.R untouched).  Everything else in R/ECS.R keeps working unchanged:
# element_sim / element_sim_elscore reach the GPU through disjointECS / disjointECSaverage (same names, same signatures),
# merge_partitions / element_consistency / element_agreement / element_sim_matrix through the wrappers below.
#
# foreach %dopar% (R/ECS.R:953) is bypassed on purpose: PSOCK workers are separate processes that would each create their
# own CUDA contexts; one process drives all selected GPUs instead (ca_init with several devices, see INTEGRATION.md).

# A membership vector as the integer labels the library takes.  Integers and factors pass through (factor codes),
# doubles are truncated by Rcpp exactly as for the reference's disjointECS (src/RcppExports.cpp:34-35); character vectors
# -- which the reference sends down a slow R loop (R/ECS.R:184-185, 468-507) -- are recoded to factor codes: ECS is
# invariant to a relabelling of either partition.
ecs_b200_labels <- function(mb) {
    if (is.character(mb)) {
        return(as.integer(factor(mb)))
    }
    mb
}

# calculate the similarity matrix when all the objects are flat disjoint memberships
element_sim_matrix_flat_disjoint <- function(mb_list, alpha = 0.9, output_type = "matrix") {
    if (!(output_type %in% c("data.frame", "matrix"))) {
        stop("output_type must be data.frame or matrix.")
    }

    n_clusterings <- length(mb_list)

    if (n_clusterings == 1) {
        return(matrix(1, nrow = 1, ncol = 1))
    }

    # all n (n - 1) / 2 pairs in one call: symmetric matrix of mean(ECS), unit diagonal
    sim_matrix <- ecsSimMatrix(lapply(unname(mb_list), ecs_b200_labels))

    if (output_type == "data.frame") {
        sim_matrix <- data.frame(
            i.clustering = rep(1:n_clusterings, n_clusterings),
            j.clustering = rep(1:n_clusterings, each = n_clusterings),
            element_sim = c(sim_matrix)
        )
    }

    return(sim_matrix)
}

# merge partitions that are identical (meaning the ecs threshold is 1)
merge_identical_partitions <- function(clustering_list) {
    n_partitions <- length(clustering_list)

    if (n_partitions == 1) {
        return(clustering_list)
    }

    # greedy first-match merge with the reference's decision rule (are_identical_memberships, R/ECS.R:984-1016), evaluated
    # on the device without one contingency table per comparison: $keep = indices of the kept partitions (in order of first
    # appearance), $freq = their summed frequencies
    merged <- ecsMergeIdentical(
        lapply(unname(clustering_list), function(x) ecs_b200_labels(x$mb)),
        as.numeric(sapply(clustering_list, function(x) x$freq))
    )

    merged_partitions <- unname(clustering_list[merged$keep])
    for (i in seq_along(merged_partitions)) {
        merged_partitions[[i]]$freq <- merged$freq[i]
    }

    merged_partitions
}

weighted_element_consistency <- function(clustering_list,
                                         weights = NULL,
                                         calculate_sim_matrix = FALSE) {
    n_clusterings <- length(clustering_list)

    if (n_clusterings <= 1) {
        stop("At least two clusterings are required to calculate the consistency.")
    }

    n_points <- length(clustering_list[[1]])
    for (mb_vector in clustering_list) {
        if (length(mb_vector) != n_points) {
            stop("clustering1 and clustering2 do not have the same length.")
        }
    }

    # ecc[n] = [ sum_i w_i (w_i - 1) / 2 + sum_{i < j} w_i w_j ECS_ij[n] ] / [ W (W - 1) / 2 ], all pairs in one call
    result <- ecsConsistency(
        lapply(unname(clustering_list), ecs_b200_labels),
        if (is.null(weights)) NULL else as.numeric(weights),
        calculate_sim_matrix
    )

    consistency <- result$ecc
    # element_sim_elscore hands the names of its first argument to the scores (R/ECS.R:193)
    names(consistency) <- names(clustering_list[[1]])

    if (!calculate_sim_matrix) {
        return(consistency)
    }

    return(list(ecc = consistency, ecs_matrix = result$ecs_matrix))
}

# calculate the element agreement when the clustering list and the reference
# are flat disjoint membership vectors
element_agreement_flat_disjoint <- function(reference_clustering,
                                            clustering_list,
                                            alpha = 0.9) {
    n_points <- length(reference_clustering)
    for (mb_vector in clustering_list) {
        if (length(mb_vector) != n_points) {
            stop("The partitions do not have the same length")
        }
    }

    # mean over the list of ECS(reference, member), one call
    avg_agreement <- ecsAgreement(
        as.integer(ecs_b200_labels(reference_clustering)),
        lapply(unname(clustering_list), ecs_b200_labels)
    )

    return(avg_agreement)
}
