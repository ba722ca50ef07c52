// Header-only reader/writer of vif's COLUMN-ORIENTED FITS tables for builds without cfitsio.
//
// The reference reads and writes its catalogs through cfitsio (io/fits/table.hpp:749,1047-1086,
// 1280-1310; io/fits/base.hpp:56-57,148-152,283-287). cfitsio is an external C library; where it is
// absent the reference is compiled with -DNO_CFITSIO and `vif::fits` does not exist at all
// (io/fits.hpp:10), so the qxmatch2 front end (tools/qxmatch2/qxmatch2.cpp:112-113,154) and
// catalog_merge (astro/catalog_merge.hpp:182,197,202) cannot be built. This header supplies exactly
// the `vif::fits::read_table` / `write_table` overloads those two callers use, written against the
// FITS standard itself (2880-byte blocks, 80-character cards, big-endian data):
//   one BINTABLE extension with ONE row; each column is a vector cell, TFORMn = '<count><type>';
//   columns of more than one dimension carry TDIMn = '(fastest,...,slowest)'; names are upper case.
// Types: D (double), E (float), K (64-bit integers; vif maps uint_t to K with the same bits, so
// npos reads back as -1 in signed readers), J (32-bit integers), I, B. Reading converts to the
// destination's element type as the reference does. Host-side I/O only (north star: "FITS/ASCII
// I/O stays on the host"). The Python twin is vif_b200/colfits.py; tests/test_gpu_cpp_tools.py
// checks the two against each other and against the reference's own fixture files.
//
// Only compiled when NO_CFITSIO is defined: with cfitsio present the reference's own vif::fits is used.
#ifndef VIF_B200_IO_COLFITS_HPP
#define VIF_B200_IO_COLFITS_HPP
#ifdef NO_CFITSIO

#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <type_traits>
#include <vector>

namespace vif {
namespace fits {
namespace colfits_impl {
    constexpr std::size_t block = 2880;

    struct column_t {
        std::string name;
        char        code = 0;          // D E K J I B
        std::size_t count = 0, offset = 0;
        std::vector<uint_t> dims;      // slowest first (vif order)
    };

    inline std::size_t type_size(char c) {
        switch (c) { case 'D': case 'K': return 8; case 'E': case 'J': return 4; case 'I': return 2; default: return 1; }
    }

    inline std::string upper(std::string s) { for (auto& c : s) c = (char)std::toupper((unsigned char)c); return s; }
    inline std::string trim(const std::string& s) {
        std::size_t a = s.find_first_not_of(" \t"), b = s.find_last_not_of(" \t");
        return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
    }

    // "r.id, r.d , xm.rid" -> {"ID", "D", "RID"}: upper-cased last component (io/ascii.hpp:828 + table.hpp)
    inline std::vector<std::string> macro_names(const std::string& s) {
        std::vector<std::string> out;
        std::size_t p = 0;
        while (p <= s.size()) {
            std::size_t q = s.find(',', p);
            if (q == std::string::npos) q = s.size();
            std::string n = trim(s.substr(p, q - p));
            std::size_t d = n.find_last_of(".>");
            if (d != std::string::npos) n = n.substr(d + 1);
            out.push_back(upper(n));
            p = q + 1;
        }
        return out;
    }

    struct table_t {
        std::vector<char>     raw;
        std::size_t           data = 0;
        std::vector<column_t> cols;
    };

    inline bool card_value(const char* card, std::string& key, std::string& val) {
        key = trim(std::string(card, 8));
        if (card[8] != '=' ) { val.clear(); return false; }
        std::string v(card + 10, 70);
        std::size_t q = v.find('\'');
        if (q != std::string::npos && trim(v.substr(0, q)).empty()) {
            std::size_t e = v.find('\'', q + 1);
            val = trim(v.substr(q + 1, e == std::string::npos ? std::string::npos : e - q - 1));
        } else {
            val = trim(v.substr(0, v.find('/')));
        }
        return true;
    }

    // loads the first binary-table extension of `file`
    inline table_t load(const std::string& file) {
        table_t t;
        std::ifstream f(file, std::ios::binary);
        vif_check(f.is_open(), "cannot open '"+file+"'");
        t.raw.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
        std::size_t off = 0;
        while (off + block <= t.raw.size()) {
            // one header
            std::vector<std::pair<std::string,std::string>> cards;
            bool end = false;
            while (!end) {
                vif_check(off + block <= t.raw.size(), "truncated FITS header in '"+file+"'");
                for (int i = 0; i < 36 && !end; ++i) {
                    std::string k, v;
                    const char* c = t.raw.data() + off + 80*i;
                    if (std::strncmp(c, "END     ", 8) == 0) { end = true; break; }
                    if (card_value(c, k, v)) cards.push_back({k, v});
                }
                off += block;
            }
            auto get = [&](const std::string& k, const std::string& def) {
                for (auto& c : cards) if (c.first == k) return c.second;
                return def;
            };
            long naxis = std::stol(get("NAXIS", "0"));
            std::size_t size = 0;
            if (naxis > 0) {
                size = (std::size_t)std::labs(std::stol(get("BITPIX", "8")))/8;
                for (long k = 1; k <= naxis; ++k) size *= (std::size_t)std::stoll(get("NAXIS"+std::to_string(k), "0"));
                size = size*(std::size_t)std::stoll(get("GCOUNT", "1")) + (std::size_t)std::stoll(get("PCOUNT", "0"));
            }
            if (get("XTENSION", "") == "BINTABLE") {
                vif_check(get("NAXIS2", "0") == "1", "'"+file+"' is a row-oriented table; column-oriented expected");
                t.data = off;
                std::size_t pos = 0;
                long nf = std::stol(get("TFIELDS", "0"));
                for (long k = 1; k <= nf; ++k) {
                    column_t c;
                    std::string tf = get("TFORM"+std::to_string(k), "");
                    std::size_t i = 0;
                    while (i < tf.size() && std::isdigit((unsigned char)tf[i])) ++i;
                    c.count = i ? (std::size_t)std::stoull(tf.substr(0, i)) : 1;
                    c.code = i < tf.size() ? tf[i] : '?';
                    c.name = upper(get("TTYPE"+std::to_string(k), "COL"+std::to_string(k)));
                    c.offset = pos;
                    std::string td = get("TDIM"+std::to_string(k), "");
                    if (!td.empty()) {
                        std::size_t a = td.find('('), b = td.find(')');
                        std::string in = td.substr(a + 1, b - a - 1);
                        std::vector<uint_t> d;
                        std::size_t p = 0;
                        while (p <= in.size()) {
                            std::size_t q = in.find(',', p);
                            if (q == std::string::npos) q = in.size();
                            d.push_back((uint_t)std::stoull(trim(in.substr(p, q - p))));
                            p = q + 1;
                        }
                        c.dims.assign(d.rbegin(), d.rend());          // FITS: fastest first; vif: slowest first
                    } else {
                        c.dims = {c.count};
                    }
                    pos += c.count*type_size(c.code);
                    t.cols.push_back(c);
                }
                vif_check(t.data + pos <= t.raw.size(), "truncated table data in '"+file+"'");
                return t;
            }
            off += (size + block - 1)/block*block;
        }
        vif_check(false, "'"+file+"' holds no binary table extension");
        return t;
    }

    template<typename T>
    inline T load_be(const char* p) {
        unsigned char b[sizeof(T)];
        for (std::size_t i = 0; i < sizeof(T); ++i) b[i] = (unsigned char)p[sizeof(T) - 1 - i];
        T v;
        std::memcpy(&v, b, sizeof(T));
        return v;
    }

    template<typename T>
    inline T element(const table_t& t, const column_t& c, std::size_t i) {
        const char* p = t.raw.data() + t.data + c.offset + i*type_size(c.code);
        switch (c.code) {
            case 'D': return (T)load_be<double>(p);
            case 'E': return (T)load_be<float>(p);
            case 'K': return (T)load_be<std::int64_t>(p);
            case 'J': return (T)load_be<std::int32_t>(p);
            case 'I': return (T)load_be<std::int16_t>(p);
            default:  return (T)(unsigned char)*p;
        }
    }

    inline const column_t& find(const table_t& t, const std::string& file, const std::string& name) {
        const std::string n = upper(name);
        for (auto& c : t.cols) if (c.name == n) return c;
        vif_check(false, "cannot find column '"+n+"' in '"+file+"'");
        return t.cols.front();
    }

    // vectors take the column's dimensions (they must agree in number), scalars need a single element
    template<std::size_t D, typename T>
    inline void fetch(const table_t& t, const std::string& file, const std::string& name, vec<D,T>& v) {
        const column_t& c = find(t, file, name);
        vif_check(c.dims.size() == D, "column '"+c.name+"' of '"+file+"' has "+to_string(c.dims.size())+
            " dimensions, "+to_string(D)+" expected");
        for (std::size_t k = 0; k < D; ++k) v.dims[k] = c.dims[k];
        v.resize();
        for (std::size_t i = 0; i < c.count; ++i) v.data[i] = element<typename vec<D,T>::dtype>(t, c, i);
    }
    template<typename T, typename enable = typename std::enable_if<std::is_arithmetic<T>::value>::type>
    inline void fetch(const table_t& t, const std::string& file, const std::string& name, T& v) {
        const column_t& c = find(t, file, name);
        vif_check(c.count == 1, "column '"+c.name+"' of '"+file+"' is not a scalar");
        v = element<T>(t, c, 0);
    }

    inline void fetch_pairs(const table_t&, const std::string&) {}
    template<typename V, typename... Args>
    inline void fetch_pairs(const table_t& t, const std::string& file, const std::string& name, V& v, Args&&... rest) {
        fetch(t, file, name, v);
        fetch_pairs(t, file, std::forward<Args>(rest)...);
    }

    inline void fetch_list(const table_t&, const std::string&, const std::vector<std::string>&, std::size_t) {}
    template<typename V, typename... Args>
    inline void fetch_list(const table_t& t, const std::string& file, const std::vector<std::string>& names, std::size_t k,
        V& v, Args&... rest) {
        fetch(t, file, names[k], v);
        fetch_list(t, file, names, k + 1, rest...);
    }

    // ---- writing ----
    template<typename T> struct code_of;
    template<> struct code_of<double>        { static constexpr char value = 'D'; };
    template<> struct code_of<float>         { static constexpr char value = 'E'; };
    template<> struct code_of<std::int64_t>  { static constexpr char value = 'K'; };
    template<> struct code_of<std::uint64_t> { static constexpr char value = 'K'; };
    template<> struct code_of<long long>     { static constexpr char value = 'K'; };
    template<> struct code_of<unsigned long long> { static constexpr char value = 'K'; };
    template<> struct code_of<std::int32_t>  { static constexpr char value = 'J'; };
    template<> struct code_of<std::uint32_t> { static constexpr char value = 'J'; };
    template<> struct code_of<std::int16_t>  { static constexpr char value = 'I'; };
    template<> struct code_of<unsigned char> { static constexpr char value = 'B'; };
    template<> struct code_of<char>          { static constexpr char value = 'B'; };
    template<> struct code_of<bool>          { static constexpr char value = 'B'; };

    struct out_column_t {
        std::string name;
        char code;
        std::vector<uint_t> dims;
        std::string bytes;             // big-endian payload
    };

    template<typename T>
    inline void push_be(std::string& s, T v) {
        unsigned char b[sizeof(T)];
        std::memcpy(b, &v, sizeof(T));
        for (std::size_t i = 0; i < sizeof(T); ++i) s.push_back((char)b[sizeof(T) - 1 - i]);
    }

    template<std::size_t D, typename T>
    inline void stage(std::vector<out_column_t>& cols, const std::string& name, const vec<D,T>& v) {
        using dt = typename std::remove_cv<typename vec<D,T>::dtype>::type;
        if (v.size() == 0) return;                        // cfitsio (and the reference) skip empty columns
        out_column_t c;
        c.name = upper(name); c.code = code_of<dt>::value;
        for (std::size_t k = 0; k < D; ++k) c.dims.push_back(v.dims[k]);
        c.bytes.reserve(v.size()*sizeof(dt));
        for (uint_t i = 0; i < v.size(); ++i) push_be<dt>(c.bytes, (dt)v.safe[i]);
        cols.push_back(std::move(c));
    }
    template<typename T, typename enable = typename std::enable_if<std::is_arithmetic<T>::value>::type>
    inline void stage(std::vector<out_column_t>& cols, const std::string& name, const T& v) {
        out_column_t c;
        c.name = upper(name); c.code = code_of<typename std::remove_cv<T>::type>::value;
        c.dims = {1};
        push_be<T>(c.bytes, v);
        cols.push_back(std::move(c));
    }

    inline void stage_pairs(std::vector<out_column_t>&) {}
    template<typename V, typename... Args>
    inline void stage_pairs(std::vector<out_column_t>& cols, const std::string& name, const V& v, const Args&... rest) {
        stage(cols, name, v);
        stage_pairs(cols, rest...);
    }
    inline void stage_list(std::vector<out_column_t>&, const std::vector<std::string>&, std::size_t) {}
    template<typename V, typename... Args>
    inline void stage_list(std::vector<out_column_t>& cols, const std::vector<std::string>& names, std::size_t k, const V& v,
        const Args&... rest) {
        stage(cols, names[k], v);
        stage_list(cols, names, k + 1, rest...);
    }

    inline std::string card(const std::string& key, const std::string& value, bool quoted, const std::string& comment = "") {
        char buf[128];
        std::string v = value;
        if (quoted) {
            std::string q = "'"+value;
            while (q.size() < 9) q.push_back(' ');
            q.push_back('\'');
            std::snprintf(buf, sizeof(buf), "%-8s= %-20s", key.c_str(), q.c_str());
        } else {
            std::snprintf(buf, sizeof(buf), "%-8s= %20s", key.c_str(), v.c_str());
        }
        std::string s = buf;
        if (!comment.empty()) s += " / "+comment;
        s.resize(80, ' ');
        return s;
    }

    inline void pad(std::string& s, char fill) { while (s.size() % block) s.push_back(fill); }

    inline void save(const std::string& file, const std::vector<out_column_t>& cols) {
        std::string prim = card("SIMPLE", "T", false, "file does conform to FITS standard")+card("BITPIX", "8", false)+
            card("NAXIS", "0", false)+card("EXTEND", "T", false);
        prim += std::string("END")+std::string(77, ' ');
        pad(prim, ' ');
        std::size_t width = 0;
        for (auto& c : cols) width += c.bytes.size();
        std::string hdr = card("XTENSION", "BINTABLE", true, "binary table extension")+card("BITPIX", "8", false)+
            card("NAXIS", "2", false)+card("NAXIS1", std::to_string(width), false, "width of table in bytes")+
            card("NAXIS2", "1", false, "number of rows in table")+card("PCOUNT", "0", false)+card("GCOUNT", "1", false)+
            card("TFIELDS", std::to_string(cols.size()), false)+card("EXTNAME", "RESULT_BINARY", true);
        std::size_t k = 1;
        for (auto& c : cols) {
            std::size_t n = c.bytes.size()/type_size(c.code);
            hdr += card("TTYPE"+std::to_string(k), c.name, true, "label for field");
            hdr += card("TFORM"+std::to_string(k), std::to_string(n)+std::string(1, c.code), true, "format of field");
            if (c.dims.size() > 1) {
                std::string d = "(";
                for (std::size_t i = c.dims.size(); i-- > 0;) { d += std::to_string(c.dims[i]); if (i) d += ","; }
                hdr += card("TDIM"+std::to_string(k), d+")", true);
            }
            ++k;
        }
        hdr += std::string("END")+std::string(77, ' ');
        pad(hdr, ' ');
        std::string data;
        data.reserve(width + block);
        for (auto& c : cols) data += c.bytes;
        pad(data, '\0');
        std::ofstream f(file, std::ios::binary | std::ios::trunc);
        vif_check(f.is_open(), "cannot write '"+file+"'");
        f.write(prim.data(), (std::streamsize)prim.size());
        f.write(hdr.data(), (std::streamsize)hdr.size());
        f.write(data.data(), (std::streamsize)data.size());
    }
}

    // fits::read_table(file, "name1", v1, "name2", v2, ...)          (qxmatch2.cpp:112-113)
    template<typename V, typename... Args>
    void read_table(const std::string& file, const std::string& name, V& v, Args&&... rest) {
        colfits_impl::table_t t = colfits_impl::load(file);
        colfits_impl::fetch_pairs(t, file, name, v, std::forward<Args>(rest)...);
    }
    template<typename V, typename... Args>
    void read_table(const std::string& file, const char* name, V& v, Args&&... rest) {
        read_table(file, std::string(name), v, std::forward<Args>(rest)...);
    }

    // fits::read_table(file, ftable(a, b.c, ...))                       (qxmatch.hpp:26, catalog_merge.hpp:182,202)
    template<typename... Args>
    void read_table(const std::string& file, impl::ascii_impl::macroed_t, const std::string& names, Args&... args) {
        colfits_impl::table_t t = colfits_impl::load(file);
        colfits_impl::fetch_list(t, file, colfits_impl::macro_names(names), 0, args...);
    }

    // fits::write_table(file, "name1", v1, ...) and fits::write_table(file, ftable(...))   (qxmatch.hpp:21, catalog_merge.hpp:197)
    template<typename V, typename... Args>
    void write_table(const std::string& file, const std::string& name, const V& v, const Args&... rest) {
        std::vector<colfits_impl::out_column_t> cols;
        colfits_impl::stage_pairs(cols, name, v, rest...);
        colfits_impl::save(file, cols);
    }
    template<typename... Args>
    void write_table(const std::string& file, impl::ascii_impl::macroed_t, const std::string& names, const Args&... args) {
        std::vector<colfits_impl::out_column_t> cols;
        colfits_impl::stage_list(cols, colfits_impl::macro_names(names), 0, args...);
        colfits_impl::save(file, cols);
    }
}
}

#endif
#endif
