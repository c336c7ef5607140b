"""Tabulate the per-kernel accounting of every bench JSON saved under profiles/."""
import json, glob, pathlib, sys
root = pathlib.Path(__file__).resolve().parents[1]
files = sorted(glob.glob(str(root / "profiles" / "r*_bench256_*.json")), key=lambda f: pathlib.Path(f).stat().st_mtime)
rows = []
for f in files:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    k = d.get('kernels', {})
    def ms(prefix):
        return sum(v['ms_per_step'] for n, v in k.items() if n.startswith(prefix))
    rows.append((pathlib.Path(f).name, d['value'], d['ms_per_step'], d['e2e']['value'] if d.get('e2e') else None,
                 ms('transform'), ms('pencil_solve'), ms('pointwise'), ms('pencil_matvec'), ms('pencil_gather') + ms('pencil_scatter'),
                 k.get('pencil_solve', {}).get('gbps'), d.get('gpu_launches')))
print("| run | steps/s | ms/step | e2e steps/s | transforms ms | solve ms | pointwise ms | matvec ms | gather+scatter ms | solve GB/s | launches |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print("| " + " | ".join(f"{x:.2f}" if isinstance(x, float) else str(x) for x in r) + " |")
