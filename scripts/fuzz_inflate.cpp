// build + run (host only): g++ -O1 -g -std=c++17 -fsanitize=address,undefined scripts/fuzz_inflate.cpp robosat_b200/csrc/rsb_inflate.cpp -lz -o /tmp/fuzz_inflate && /tmp/fuzz_inflate 1
// ASAN/UBSAN fuzz of the library inflate: corrupted / truncated zlib streams in buffers with EXACTLY the documented padding
#include "../robosat_b200/csrc/rsb_inflate.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <zlib.h>
using namespace rsb;
int main(int argc,char**argv){
  unsigned seed=argc>1?atoi(argv[1]):1; srand(seed);
  long ok=0,rej=0,wrong=0;
  for(int it=0; it<6000; ++it){
    size_t n = (it%7==0)? rand()%200000 : rand()%5000;
    std::vector<uint8_t> src(n);
    int mode=it%5;
    for(size_t i=0;i<n;++i) src[i]= mode==0? rand() : mode==1? (rand()%4) : mode==2? (uint8_t)(i/7) : mode==3? 0 : (uint8_t)("abcabcabd"[i%9]);
    int level = it%10; int strat = (it/3)%5; // default, filtered, huffman, rle, fixed
    z_stream zs; memset(&zs,0,sizeof zs); deflateInit2(&zs, level, Z_DEFLATED, 15, 8, strat);
    std::vector<uint8_t> comp(compressBound(n)+64);
    zs.next_in=src.data(); zs.avail_in=n; zs.next_out=comp.data(); zs.avail_out=comp.size(); deflate(&zs,Z_FINISH); size_t cn=zs.total_out; deflateEnd(&zs);
    for(int variant=0; variant<4; ++variant){
      size_t in_len=cn; std::vector<uint8_t> in(comp.begin(), comp.begin()+cn);
      size_t out_len=n;
      if(variant==1 && cn>2){ int flips=1+rand()%3; for(int f=0;f<flips;++f) in[rand()%cn]^=1<<(rand()%8); }
      if(variant==2 && cn>1){ in_len=rand()%cn; in.resize(in_len); }
      if(variant==3){ out_len = n? (rand()%(2*n+1)) : 1; }
      // exact padding, allocated on the heap so ASAN sees any access beyond it
      uint8_t* ib=(uint8_t*)malloc(in_len+kInflateInPad); if (in_len) memcpy(ib, in.data(), in_len); memset(ib+in_len,0,kInflateInPad);
      uint8_t* ob=(uint8_t*)malloc(out_len+kInflateOutPad);
      int rc = in_len>=6 ? rsb_inflate_zlib_padded(ib,in_len,ob,out_len) : -2;
      if(rc==0){ if(out_len==n && memcmp(ob,src.data(),n)==0) ok++; else { // a successful decode must be THE data (adler verified) unless sizes differ legitimately
           if(variant==0||variant==3) { if(!(out_len==n)) wrong++; else wrong++; } else { /* corrupted but still adler-valid and same bytes? */ if(out_len!=n||memcmp(ob,src.data(),n)) wrong++; else ok++; } } }
      else rej++;
      free(ib); free(ob);
    }
  }
  printf("seed %u ok %ld rejected %ld wrong %ld\n",seed,ok,rej,wrong);
  return wrong?1:0;
}
