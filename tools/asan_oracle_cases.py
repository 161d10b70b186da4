import ctypes as C, json, random
O=C.CDLL('/tmp/fourmc_asan/liboracle_asan.so')
O.fmo_zstd_decompress.restype=C.c_longlong
O.fmo_zstd_decompress.argtypes=[C.c_char_p,C.c_longlong,C.c_char_p,C.c_longlong]
O.fmo_lz4_decompress_safe.restype=C.c_int
O.fmo_lz4_decompress_safe.argtypes=[C.c_char_p,C.c_char_p,C.c_int,C.c_int]
O.fmo_4mz_decompress.restype=C.c_longlong
O.fmo_4mz_decompress.argtypes=[C.c_char_p,C.c_size_t,C.c_char_p,C.c_size_t]
O.fmo_4mc_decompress.restype=C.c_longlong
O.fmo_4mc_decompress.argtypes=[C.c_char_p,C.c_size_t,C.c_char_p,C.c_size_t]
n=0
for e in json.load(open('tests/golden/zstd_decode.json')):
    src=bytes.fromhex(e['hex'])
    for cap,ret,x in e['runs']:
        out=C.create_string_buffer(max(cap,1))          # exact capacity: overruns are caught
        O.fmo_zstd_decompress(out,cap,src,len(src)); n+=1
for e in json.load(open('tests/golden/lz4_decode.json')):
    src=bytes.fromhex(e['hex']); cap=e['cap']
    out=C.create_string_buffer(max(cap,1))
    O.fmo_lz4_decompress_safe(src,out,len(src),cap); n+=1
rng=random.Random(3)
for name,fn in (('logtext_128k.z1.4mz',O.fmo_4mz_decompress),('logtext_128k.z3.4mz',O.fmo_4mz_decompress),('logtext_128k.l1.4mc',O.fmo_4mc_decompress),('logtext_128k.l3.4mc',O.fmo_4mc_decompress)):
    good=open('tests/golden/'+name,'rb').read()
    for t in range(400):
        m=bytearray(good)
        for _ in range(rng.randrange(1,4)):
            k=rng.randrange(3); at=rng.randrange(len(m))
            if k==0: m[at]^=1<<rng.randrange(8)
            elif k==1: del m[at:at+rng.randrange(1,50)]
            else: m[at:at]=rng.randbytes(rng.randrange(1,8))
        out=C.create_string_buffer(131072)
        fn(bytes(m),len(m),out,131072); n+=1
print("ran",n,"cases clean")
