// TEST INFRASTRUCTURE ONLY: prints the libstdc++ unordered_map<size_t,_> bucket-count ladder
// (element count at which each rehash happens, new bucket count).  Used to pin k_ladder in neighbors.c.
#include <unordered_map>
#include <cstdio>
#include <cstddef>
int main(){
  std::unordered_map<std::size_t,int> m;
  size_t last = m.bucket_count();
  printf("init %zu\n", last);
  for (size_t i=0;i<60000000ull;i++){ m.emplace(i*7919+3,0); if(m.bucket_count()!=last){ last=m.bucket_count(); printf("n=%zu nb=%zu\n", m.size(), last);} }
}
