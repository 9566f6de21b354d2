#!/usr/bin/env python
"""r/apply_ecs_b200.py PATH/TO/ClustAssess [--check] -- splice the B200 wrappers of r/ECS_b200.R into a ClustAssess checkout.

Rewrites exactly four function definitions of R/ECS.R -- element_sim_matrix_flat_disjoint (R/ECS.R:931-981),
merge_identical_partitions (:1021-1054), weighted_element_consistency (:1446-1500), element_agreement_flat_disjoint
(:1625-1648) -- with the ones of r/ECS_b200.R, puts the helper ecs_b200_labels in front of the first of them, and leaves
every other line (roxygen blocks included) untouched.  --check only reports what would change.  The C++ side is
r/calculate_ecs_b200.cpp (copy it over src/calculate_ecs.cpp, then Rcpp::compileAttributes()); see INTEGRATION.md.
"""
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
WRAPPERS = ["element_sim_matrix_flat_disjoint", "merge_identical_partitions", "weighted_element_consistency",
            "element_agreement_flat_disjoint"]
HELPER = "ecs_b200_labels"


def function_span(text, name):
    """(start, end) character offsets of `name <- function(...) { ... }` at top level; braces inside strings and comments
    are skipped."""
    m = re.search(r"^%s\s*<-\s*function\s*\(" % re.escape(name), text, flags=re.M)
    if m is None:
        raise SystemExit("definition of %s not found" % name)
    i, depth, seen_body = m.end(), 0, False
    paren = 1
    quote = None
    while i < len(text):
        ch = text[i]
        if quote:
            if ch == "\\":
                i += 1
            elif ch == quote:
                quote = None
        elif ch in "\"'`":
            quote = ch
        elif ch == "#":
            while i < len(text) and text[i] != "\n":
                i += 1
        elif paren > 0:  # still inside the formal argument list
            paren += ch == "("
            paren -= ch == ")"
        elif ch == "{":
            depth += 1
            seen_body = True
        elif ch == "}":
            depth -= 1
            if seen_body and depth == 0:
                return m.start(), i + 1
        i += 1
    raise SystemExit("unbalanced braces in the definition of %s" % name)


def splice(ecs_r_text, wrappers_text):
    out = ecs_r_text
    first = min(function_span(out, n)[0] for n in WRAPPERS)
    changed = []
    for name in WRAPPERS:
        a, b = function_span(out, name)
        na, nb = function_span(wrappers_text, name)
        new = wrappers_text[na:nb]
        if out[a:b] != new:
            changed.append(name)
        out = out[:a] + new + out[b:]
    if re.search(r"^%s\s*<-\s*function" % HELPER, out, flags=re.M) is None:
        ha, hb = function_span(wrappers_text, HELPER)
        # its comment block travels with it
        nl = "\r\n" if "\r\n" in wrappers_text else "\n"
        block_start = wrappers_text.rfind(nl + nl, 0, ha) + 2 * len(nl)
        out = out[:first] + wrappers_text[block_start:hb] + nl + nl + out[first:]
        changed.append(HELPER)
    return out, changed


def main(argv):
    args = [a for a in argv if not a.startswith("--")]
    if len(args) != 1:
        raise SystemExit(__doc__)
    path = os.path.join(args[0], "R", "ECS.R")
    text = open(path, newline="").read()  # keep the checkout's line endings (the CRAN sources use CRLF)
    wrappers = open(os.path.join(HERE, "ECS_b200.R")).read()
    if "\r\n" in text:
        wrappers = wrappers.replace("\n", "\r\n")
    new, changed = splice(text, wrappers)
    print("R/ECS.R: %s" % (", ".join(changed) if changed else "already patched"))
    if "--check" not in argv and new != text:
        open(path, "w", newline="").write(new)
        print("written: %s" % path)


if __name__ == "__main__":
    main(sys.argv[1:])
