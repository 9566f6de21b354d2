"""The R side of the drop-in: r/ECS_b200.R holds the four set-level wrappers of R/ECS.R rewritten to make one
.Call each, r/apply_ecs_b200.py splices them into a ClustAssess checkout.  No R interpreter exists in the build container, so
what can be pinned here is the splice itself: exactly the four definitions change (reference lines R/ECS.R:931-981, 1021-1054,
1446-1500, 1625-1648), nothing else, idempotently -- on a synthetic file always, on the real R/ECS.R when /root/reference is
present -- and that every routine the wrappers call is exported by r/calculate_ecs_b200.cpp with matching arity."""
import importlib.util
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("apply_ecs_b200", os.path.join(ROOT, "r", "apply_ecs_b200.py"))
ap = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ap)
WRAPPERS = open(os.path.join(ROOT, "r", "ECS_b200.R")).read()


def test_function_span_skips_braces_in_strings_and_comments():
    text = 'a <- 1\nf <- function(x, y = "}") {\n  # a } in a comment\n  if (x) { "{" } else { g(function(z) { z }) }\n}\nb <- 2\n'
    a, b = ap.function_span(text, "f")
    assert text[a:b].startswith("f <- function") and text[b:] == "\nb <- 2\n"


def test_splice_replaces_exactly_the_four_wrappers_and_is_idempotent():
    stub = "\n\n".join("%s <- function(a, b = 0.9) {\n    old_body_%d(a) # {\n}" % (n, k) for k, n in enumerate(ap.WRAPPERS))
    text = "keep_me <- function() { 1 }\n\n#' roxygen stays\n" + stub + "\n\ntail_stays <- 3\n"
    new, changed = ap.splice(text, WRAPPERS)
    assert set(changed) == set(ap.WRAPPERS) | {ap.HELPER}
    assert new.startswith("keep_me <- function() { 1 }\n\n#' roxygen stays\n") and new.endswith("\n\ntail_stays <- 3\n")
    assert "old_body_" not in new and new.count("ecs_b200_labels <- function") == 1
    again, changed2 = ap.splice(new, WRAPPERS)
    assert again == new and changed2 == []


def test_wrappers_call_only_routines_the_glue_exports():
    glue = open(os.path.join(ROOT, "r", "calculate_ecs_b200.cpp")).read()
    exported = {}
    for chunk in glue.split("// [[Rcpp::export]]\n")[1:]:
        sig = chunk.split("{")[0]
        name = sig.split("(")[0].split()[-1]
        exported[name] = len([a for a in sig[sig.index("(") + 1:sig.rindex(")")].split(",") if a.strip()])
    calls = {"ecsSimMatrix": 1, "ecsMergeIdentical": 2, "ecsConsistency": 3, "ecsAgreement": 2}
    for name, nargs in calls.items():
        assert exported.get(name) == nargs, name
        assert re.search(r"\b%s\(" % name, WRAPPERS), name


@pytest.mark.skipif(not os.path.exists("/root/reference/R/ECS.R"), reason="the reference checkout is only present in the build container")
def test_splice_into_the_reference_file_touches_only_the_four_definitions(tmp_path):
    import difflib

    src = open("/root/reference/R/ECS.R", newline="").read()
    wr = WRAPPERS.replace("\n", "\r\n") if "\r\n" in src else WRAPPERS
    new, changed = ap.splice(src, wr)
    assert set(changed) == set(ap.WRAPPERS) | {ap.HELPER}
    old_lines, new_lines = src.splitlines(), new.splitlines()
    touched = set()
    for tag, i1, i2, _, _ in difflib.SequenceMatcher(None, old_lines, new_lines, autojunk=False).get_opcodes():
        if tag != "equal":
            touched.update(range(i1 + 1, max(i2, i1 + 1) + 1))
    allowed = set(range(929, 983)) | set(range(1019, 1056)) | set(range(1444, 1502)) | set(range(1623, 1650))
    assert touched and touched <= allowed, sorted(touched - allowed)[:10]
    for name in ("element_sim_elscore", "merge_partitions_ecs", "merge_partitions", "element_consistency", "are_identical_memberships"):
        a, b = ap.function_span(src, name)
        a2, b2 = ap.function_span(new, name)
        assert src[a:b] == new[a2:b2]  # the callers and the untouched helpers are byte-identical
