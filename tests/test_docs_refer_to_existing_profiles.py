"""Every measurement file the documents cite under profiles/ exists (brace lists like r02_bench_{1,2}gpu.json are expanded)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"`((?:profiles/)?r0[12]_[A-Za-z0-9_{},.\-]+\.(?:json|txt|log|csv|gz))`")


def expand(s):
    m = re.search(r"\{([^{}]*)\}", s)
    if not m:
        return [s]
    out = []
    for alt in m.group(1).split(","):
        out += expand(s[:m.start()] + alt + s[m.end():])
    return out


def test_cited_profiles_exist():
    missing = []
    cited = 0
    for doc in ("DESIGN.md", "README.md", "INTEGRATION.md", os.path.join("profiles", "README.md")):
        text = open(os.path.join(ROOT, doc)).read()
        for m in PAT.finditer(text):
            name = m.group(1)
            for p in expand(name if name.startswith("profiles/") else "profiles/" + name):
                cited += 1
                if not os.path.exists(os.path.join(ROOT, p)):
                    missing.append((doc, p))
    assert cited > 50
    assert not missing, missing
