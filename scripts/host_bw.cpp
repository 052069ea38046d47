#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <cstdlib>
int main(){ size_t N=size_t(384)<<20; char*a=(char*)aligned_alloc(1<<21,N),*b=(char*)aligned_alloc(1<<21,N); memset(a,1,N); memset(b,2,N);
 for(int w: {1,2,4,8}){ for(int rep=0;rep<3;++rep){ auto t0=std::chrono::steady_clock::now(); std::vector<std::thread> p; for(int i=0;i<w;++i) p.emplace_back([&,i]{ memcpy(b+N*i/w,a+N*i/w,N/w);}); for(auto&t:p)t.join(); double s=std::chrono::duration<double>(std::chrono::steady_clock::now()-t0).count(); if(rep==2) printf("threads %d: %.1f ms %.1f GB/s\n",w,s*1e3,N/s/1e9);} } }
