"""Development aid: how much of a source file is covered by 12-token n-grams that also occur in the
reference's sources (comments and whitespace stripped) -- the measure VERDICT r1 used for copy findings.
usage: python tools/ngram_check.py FILE [REFDIR]"""
import re
import sys
import glob
import os

TOK = re.compile(r"[A-Za-z_][A-Za-z_0-9]*|\d+\.?\d*(?:[eE][-+]?\d+)?|==|!=|<=|>=|->|\+\+|--|&&|\|\||<<|>>|[-+*/%=<>!&|^~?:;,.(){}\[\]#]|\"(?:\\.|[^\"\\])*\"|'(?:\\.|[^'\\])*'")


def tokens(path):
    s = open(path, errors="replace").read()
    s = re.sub(r"/\*.*?\*/", " ", s, flags=re.S)
    s = re.sub(r"//[^\n]*", " ", s)
    return TOK.findall(s)


def main():
    f = sys.argv[1]
    refdir = sys.argv[2] if len(sys.argv) > 2 else "/root/reference/lib"
    n = 12
    mine = tokens(f)
    total = {}
    for ref in sorted(glob.glob(os.path.join(refdir, "*.[ch]"))):
        rt = tokens(ref)
        grams = {tuple(rt[i:i + n]) for i in range(len(rt) - n + 1)}
        covered = [False] * len(mine)
        for i in range(len(mine) - n + 1):
            if tuple(mine[i:i + n]) in grams:
                for j in range(i, i + n):
                    covered[j] = True
        c = sum(covered)
        if c:
            total[os.path.basename(ref)] = c / max(1, len(mine))
    allg = set()
    for ref in glob.glob(os.path.join(refdir, "*.[ch]")):
        rt = tokens(ref)
        allg |= {tuple(rt[i:i + n]) for i in range(len(rt) - n + 1)}
    covered = [False] * len(mine)
    for i in range(len(mine) - n + 1):
        if tuple(mine[i:i + n]) in allg:
            for j in range(i, i + n):
                covered[j] = True
    print(f"{f}: {len(mine)} tokens, {100 * sum(covered) / max(1, len(mine)):.1f}% covered by reference 12-grams")
    for k, v in sorted(total.items(), key=lambda kv: -kv[1])[:6]:
        print(f"   {k}: {100 * v:.1f}%")


if __name__ == "__main__":
    main()
