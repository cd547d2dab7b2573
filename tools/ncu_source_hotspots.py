import csv, sys, collections
def load(fn):
    rows = []; cur_file = None; hdr = None
    for r in csv.reader(open(fn)):
        if not r: continue
        if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
        if r[0] == "Function Name": continue
        if r[0] == "Line No": hdr = r; continue
        if hdr is None: continue
        d = dict(zip(range(len(r)), r))
        rows.append((cur_file, r))
    return hdr, rows
def agg(fn):
    hdr, rows = load(fn)
    i_inst = hdr.index("Instructions Executed"); i_samp = hdr.index("# Samples"); i_thr = hdr.index("Thread Instructions Executed")
    per_line = collections.OrderedDict(); line = None
    for f, r in rows:
        if r[0] != "":   # a source line
            line = (f, int(r[0]), r[1].strip()[:90]); per_line.setdefault(line, [0, 0, 0]); continue
        try:
            per_line[line][0] += int(r[i_inst]); per_line[line][1] += int(r[i_samp]); per_line[line][2] += int(r[i_thr])
        except Exception: pass
    return per_line
for fn in sys.argv[1:]:
    pl = agg(fn)
    tot_i = sum(v[0] for v in pl.values()); tot_s = sum(v[1] for v in pl.values())
    print("==", fn, "total warp instr %.3e samples %d" % (tot_i, tot_s))
    byfile = collections.Counter(); bys = collections.Counter()
    for (f, l, s), v in pl.items(): byfile[f] += v[0]; bys[f] += v[1]
    for f, c in byfile.most_common(): print("   %-22s instr %5.1f%%  samples %5.1f%%" % (f, 100 * c / tot_i, 100 * bys[f] / tot_s))
    top = sorted(pl.items(), key=lambda kv: -kv[1][1])[:45]
    for (f, l, s), v in top:
        print("   %5.2f%% samp %5.2f%% instr  thr/inst %4.1f  %s:%d  %s" % (100 * v[1] / tot_s, 100 * v[0] / tot_i, v[2] / max(v[0], 1), f, l, s))
