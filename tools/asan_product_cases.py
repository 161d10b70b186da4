import ctypes as C, json, random
Z=C.CDLL('/tmp/fourmc_asan/zstd_shim_asan.so')
Z.zstd_shim_decompress.restype=C.c_longlong
Z.zstd_shim_decompress.argtypes=[C.c_char_p,C.c_longlong,C.c_char_p,C.c_longlong]
P=C.CDLL('/tmp/fourmc_asan/parse_shim_asan.so')
P.parse_shim.restype=C.c_int
P.parse_shim.argtypes=[C.c_char_p,C.c_int,C.c_int,C.c_void_p,C.c_void_p,C.c_void_p,C.c_int]
n=0; bad=0
for e in json.load(open('tests/golden/zstd_decode.json')):
    src=bytes.fromhex(e['hex'])
    for cap,ret,x in e['runs']:
        out=C.create_string_buffer(max(cap,1)+64)      # the shim's contract in the tests: cap + 64
        r=Z.zstd_shim_decompress(out,cap,src,len(src)); n+=1
        if (r<0)!=(ret<0) or (ret>=0 and r!=ret): bad+=1
for e in json.load(open('tests/golden/lz4_decode.json')):
    src=bytes.fromhex(e['hex'])
    # exact-size source buffer so that reads past the end are caught
    buf=(C.c_char*len(src)).from_buffer_copy(src) if src else C.create_string_buffer(1)
    r=P.parse_shim(C.cast(buf,C.c_char_p),len(src),e['cap'],None,None,None,0); n+=1
    if r!=e['ret']: bad+=1
rng=random.Random(5)
import struct
def blocks(stream):
    pos=12; out=[]
    while True:
        u,c,ck=struct.unpack(">III",stream[pos:pos+12])
        if u==0 and c==0: return out
        out.append((u,c,stream[pos+12:pos+12+c])); pos+=12+c
for name in ('logtext_128k.z1.4mz','logtext_128k.z2.4mz','logtext_128k.z4.4mz','logtext_1280k.z1.4mz'):
    for u,c,frame in blocks(open('tests/golden/'+name,'rb').read())[:2]:
        for t in range(300):
            m=bytearray(frame)
            for _ in range(rng.randrange(1,3)):
                k=rng.randrange(3); at=rng.randrange(len(m))
                if k==0: m[at]^=1<<rng.randrange(8)
                elif k==1: del m[at:at+rng.randrange(1,50)]
                else: m[at:at]=rng.randbytes(rng.randrange(1,8))
            src=bytes(m); buf=(C.c_char*len(src)).from_buffer_copy(src)
            out=C.create_string_buffer(u+64)
            Z.zstd_shim_decompress(out,u,C.cast(buf,C.c_char_p),len(src)); n+=1
print("ran",n,"cases; verdict mismatches:",bad)
