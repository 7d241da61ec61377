import subprocess, sys, re, hashlib, glob, os, json
def funcs(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    res, name, lines = {}, None, []
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            if name: res[name] = lines
            name, lines = m.group(1), []
        elif name:
            m2 = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);\s*/\*", ln)
            if m2: lines.append(m2.group(1).strip())
    if name: res[name] = lines
    return res
def norm(n):  # anonymous-namespace hash differs per translation-unit content: strip it
    return re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]+", "_GLOBAL__N_", n)
d = {}
for obj in sorted(glob.glob(os.path.join(sys.argv[1], "*.o"))):
    for n, l in funcs(obj).items():
        d[os.path.basename(obj) + ":" + norm(n)] = (len(l), hashlib.sha1("\n".join(l).encode()).hexdigest()[:16])
json.dump(d, open(sys.argv[2], "w"), indent=0, sort_keys=True)
print(len(d), "functions")
