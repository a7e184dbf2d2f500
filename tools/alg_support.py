"""Which of the reference's codes (N = 24 z, six rates) each decoding algorithm accepts: PYTHONPATH=. python tools/alg_support.py"""
import myldpccppapi_b200 as m
rates = [(0, 1, 2), (1, 2, 3), (2, 2, 3), (3, 3, 4), (4, 3, 4), (5, 5, 6)]
print("z    " + "  ".join("r%d:sp/td" % r for r, _, _ in rates))
for z in range(24, 97, 4):
    N = 24 * z
    row = []
    for r, num, den in rates:
        d = m.Decoder.wimax(N * num // den, N, r)
        res = ""
        for alg in (1, 2):
            try:
                d.set_algorithm(alg)
                res += "y"
            except m.LdpcError:
                res += "-"
        row.append(res + ":" + d.info()["path_name"][:5])
        d.close()
    print("%-4d " % z + "  ".join(row))
