// RecBole atomic `.inter` file -> remapped id columns: the on-disk format that feeds the path
// (`tests/test_data/test/test.inter` in the reference; RecBole's `Dataset` remaps the tokens and hands
// `inter_feat[uid_field]`, `inter_feat[iid_field]` to `get_norm_adj_mat`, dataset.py:60-61).
//
// Host code (no CUDA): the file is mapped, tokenised in one pass (tab-separated, header `name:type`), and the
// user / item tokens are remapped to ids in first-appearance order starting at 1 — id 0 is RecBole's [PAD].
// The id columns then go to `b200gcn_csr_from_interactions` on the device.
#include <fcntl.h>
#include <stdint.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <string_view>
#include <unordered_map>
#include <vector>

#include "b200gcn.h"

namespace b200gcn {
void set_error(const char* fmt, ...);
}

namespace {
struct InterFile {
  std::vector<int64_t> uid, iid;
  int64_t user_num = 1, item_num = 1;   // incl. [PAD]
};

struct SvHash {
  size_t operator()(std::string_view s) const noexcept { return std::hash<std::string_view>{}(s); }
};

// field `idx` (0-based) of a tab-separated line [p, e); returns false if the line has fewer fields
bool field(const char* p, const char* e, int idx, std::string_view* out) {
  const char* s = p;
  for (int k = 0;; ++k) {
    const char* t = static_cast<const char*>(memchr(s, '\t', size_t(e - s)));
    const char* fe = t ? t : e;
    if (k == idx) {
      *out = std::string_view(s, size_t(fe - s));
      return true;
    }
    if (!t) return false;
    s = t + 1;
  }
}

int header_column(const char* p, const char* e, const char* name, int fallback) {
  const size_t nl = strlen(name);
  const char* s = p;
  for (int k = 0;; ++k) {
    const char* t = static_cast<const char*>(memchr(s, '\t', size_t(e - s)));
    const char* fe = t ? t : e;
    const char* colon = static_cast<const char*>(memchr(s, ':', size_t(fe - s)));
    const size_t len = size_t((colon ? colon : fe) - s);
    if (len == nl && memcmp(s, name, nl) == 0) return k;
    if (!t) return fallback;
    s = t + 1;
  }
}
}  // namespace

extern "C" int b200gcn_inter_open(const char* path, void** handle, int64_t* n_inter, int64_t* user_num,
                                  int64_t* item_num) {
  if (!path || !handle || !n_inter || !user_num || !item_num) {
    b200gcn::set_error("b200gcn_inter_open: NULL argument");
    return B200GCN_ERR_INVALID;
  }
  const int fd = open(path, O_RDONLY);
  if (fd < 0) {
    b200gcn::set_error("cannot open %s", path);
    return B200GCN_ERR_INVALID;
  }
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size == 0) {
    close(fd);
    b200gcn::set_error("%s is empty or unreadable", path);
    return B200GCN_ERR_INVALID;
  }
  const size_t size = size_t(st.st_size);
  void* map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (map == MAP_FAILED) {
    b200gcn::set_error("mmap of %s failed", path);
    return B200GCN_ERR_INVALID;
  }
  const char* base = static_cast<const char*>(map);
  const char* end = base + size;
  auto line_end = [&](const char* p) {
    const char* nl = static_cast<const char*>(memchr(p, '\n', size_t(end - p)));
    return nl ? nl : end;
  };
  auto trim_cr = [](const char* p, const char* e) { return (e > p && e[-1] == '\r') ? e - 1 : e; };

  const char* p = base;
  const char* le = line_end(p);
  const char* he = trim_cr(p, le);
  const int ucol = header_column(p, he, "user_id", 0);
  const int icol = header_column(p, he, "item_id", 1);
  p = le < end ? le + 1 : end;

  auto* f = new InterFile();
  std::unordered_map<std::string_view, int64_t, SvHash> users, items;   // views into the mapping (kept until done)
  f->uid.reserve(size / 24);
  f->iid.reserve(size / 24);
  while (p < end) {
    le = line_end(p);
    const char* e = trim_cr(p, le);
    std::string_view u, i;
    if (e > p && field(p, e, ucol, &u) && field(p, e, icol, &i)) {
      auto iu = users.try_emplace(u, int64_t(users.size()) + 1).first;
      auto ii = items.try_emplace(i, int64_t(items.size()) + 1).first;
      f->uid.push_back(iu->second);
      f->iid.push_back(ii->second);
    }
    p = le < end ? le + 1 : end;
  }
  f->user_num = int64_t(users.size()) + 1;
  f->item_num = int64_t(items.size()) + 1;
  munmap(map, size);
  *handle = f;
  *n_inter = int64_t(f->uid.size());
  *user_num = f->user_num;
  *item_num = f->item_num;
  return B200GCN_OK;
}

extern "C" int b200gcn_inter_read(void* handle, int64_t* h_uid, int64_t* h_iid) {
  if (!handle || !h_uid || !h_iid) {
    b200gcn::set_error("b200gcn_inter_read: NULL argument");
    return B200GCN_ERR_INVALID;
  }
  auto* f = static_cast<InterFile*>(handle);
  if (!f->uid.empty()) {
    memcpy(h_uid, f->uid.data(), f->uid.size() * sizeof(int64_t));
    memcpy(h_iid, f->iid.data(), f->iid.size() * sizeof(int64_t));
  }
  return B200GCN_OK;
}

extern "C" void b200gcn_inter_close(void* handle) { delete static_cast<InterFile*>(handle); }
