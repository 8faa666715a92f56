/*
 * Small non-validating XML DOM builder implementing the subset of the libxml2 tree API
 * declared in shims/libxml/parser.h. Build-environment shim only (libxml2 is absent from
 * this image); it is not on the render hot path and holds no flam3 logic.
 *
 * Supported: XML declaration / processing instructions (skipped), comments, CDATA,
 * elements, attributes in single or double quotes, character data, the five predefined
 * entities and numeric character references. No DTDs, no namespaces, no encodings other
 * than bytes-as-is (flam3 files are ASCII/UTF-8).
 */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "parser.h"

xmlFreeFunc xmlFree = free;

/* ---------- small helpers ---------- */

static xmlChar *dup_n(const char *s, size_t n) {
   xmlChar *r = (xmlChar *)malloc(n + 1);
   if (!r) return NULL;
   memcpy(r, s, n);
   r[n] = 0;
   return r;
}

static xmlChar *dup_z(const xmlChar *s) {
   return s ? dup_n((const char *)s, strlen((const char *)s)) : NULL;
}

/* Decode entities of src[0..n) into a fresh string. */
static xmlChar *decode_text(const char *src, size_t n) {
   char *out = (char *)malloc(n + 1);
   size_t i = 0, o = 0;
   if (!out) return NULL;
   while (i < n) {
      if (src[i] != '&') {
         out[o++] = src[i++];
         continue;
      }
      /* find ';' */
      size_t j = i + 1;
      while (j < n && src[j] != ';' && j - i < 12) j++;
      if (j >= n || src[j] != ';') {
         out[o++] = src[i++];
         continue;
      }
      const char *e = src + i + 1;
      size_t el = j - i - 1;
      if (el == 3 && !strncmp(e, "amp", 3)) out[o++] = '&';
      else if (el == 2 && !strncmp(e, "lt", 2)) out[o++] = '<';
      else if (el == 2 && !strncmp(e, "gt", 2)) out[o++] = '>';
      else if (el == 4 && !strncmp(e, "quot", 4)) out[o++] = '"';
      else if (el == 4 && !strncmp(e, "apos", 4)) out[o++] = '\'';
      else if (el >= 2 && e[0] == '#') {
         unsigned long cp;
         if (e[1] == 'x' || e[1] == 'X') cp = strtoul(e + 2, NULL, 16);
         else cp = strtoul(e + 1, NULL, 10);
         /* UTF-8 encode */
         if (cp < 0x80) out[o++] = (char)cp;
         else if (cp < 0x800) {
            out[o++] = (char)(0xC0 | (cp >> 6));
            out[o++] = (char)(0x80 | (cp & 0x3F));
         } else {
            out[o++] = (char)(0xE0 | ((cp >> 12) & 0x0F));
            out[o++] = (char)(0x80 | ((cp >> 6) & 0x3F));
            out[o++] = (char)(0x80 | (cp & 0x3F));
         }
      } else {
         /* unknown entity: keep verbatim */
         memcpy(out + o, src + i, j - i + 1);
         o += j - i + 1;
      }
      i = j + 1;
   }
   out[o] = 0;
   return (xmlChar *)out;
}

static xmlNodePtr new_node(xmlElementType type, const xmlChar *name_owned) {
   xmlNodePtr n = (xmlNodePtr)calloc(1, sizeof(xmlNode));
   if (!n) return NULL;
   n->type = type;
   n->name = name_owned;
   return n;
}

static void link_child(xmlNodePtr parent, xmlNodePtr cur) {
   cur->parent = parent;
   cur->doc = parent->doc;
   cur->next = NULL;
   cur->prev = parent->last;
   if (parent->last) parent->last->next = cur;
   else parent->children = cur;
   parent->last = cur;
}

static void set_doc_rec(xmlNodePtr n, xmlDocPtr doc) {
   for (; n; n = n->next) {
      n->doc = doc;
      set_doc_rec(n->children, doc);
   }
}

static xmlNodePtr new_text(const xmlChar *content_owned) {
   xmlNodePtr t = new_node(XML_TEXT_NODE, dup_z((const xmlChar *)"text"));
   if (t) t->content = (xmlChar *)content_owned;
   return t;
}

static void free_node_list(xmlNodePtr n);

static void free_attr_list(xmlAttrPtr a) {
   while (a) {
      xmlAttrPtr nx = a->next;
      free_node_list(a->children);
      free((void *)a->name);
      free(a);
      a = nx;
   }
}

static void free_node_list(xmlNodePtr n) {
   while (n) {
      xmlNodePtr nx = n->next;
      free_node_list(n->children);
      free_attr_list(n->properties);
      free((void *)n->name);
      free(n->content);
      free(n);
      n = nx;
   }
}

/* ---------- parser ---------- */

typedef struct {
   const char *p, *end;
   int error;
} xp;

static void skip_ws(xp *x) {
   while (x->p < x->end && isspace((unsigned char)*x->p)) x->p++;
}

static int starts(xp *x, const char *lit) {
   size_t n = strlen(lit);
   return (size_t)(x->end - x->p) >= n && !strncmp(x->p, lit, n);
}

static int skip_until(xp *x, const char *lit) {
   size_t n = strlen(lit);
   while ((size_t)(x->end - x->p) >= n) {
      if (!strncmp(x->p, lit, n)) {
         x->p += n;
         return 0;
      }
      x->p++;
   }
   x->error = 1;
   return 1;
}

static int is_name_char(int c) {
   return isalnum(c) || c == '_' || c == '-' || c == '.' || c == ':' || (c & 0x80);
}

static xmlChar *parse_name(xp *x) {
   const char *s = x->p;
   while (x->p < x->end && is_name_char((unsigned char)*x->p)) x->p++;
   if (x->p == s) {
      x->error = 1;
      return NULL;
   }
   return dup_n(s, (size_t)(x->p - s));
}

static void parse_content(xp *x, xmlNodePtr parent);

static xmlNodePtr parse_element(xp *x) {
   /* at '<' followed by a name char */
   x->p++;
   xmlChar *name = parse_name(x);
   if (!name) return NULL;
   xmlNodePtr el = new_node(XML_ELEMENT_NODE, name);
   xmlAttrPtr last_attr = NULL;

   for (;;) {
      skip_ws(x);
      if (x->p >= x->end) {
         x->error = 1;
         break;
      }
      if (starts(x, "/>")) {
         x->p += 2;
         return el;
      }
      if (*x->p == '>') {
         x->p++;
         break;
      }
      /* attribute */
      xmlChar *an = parse_name(x);
      if (!an) break;
      skip_ws(x);
      if (x->p >= x->end || *x->p != '=') {
         free(an);
         x->error = 1;
         break;
      }
      x->p++;
      skip_ws(x);
      if (x->p >= x->end || (*x->p != '"' && *x->p != '\'')) {
         free(an);
         x->error = 1;
         break;
      }
      char q = *x->p++;
      const char *vs = x->p;
      while (x->p < x->end && *x->p != q) x->p++;
      if (x->p >= x->end) {
         free(an);
         x->error = 1;
         break;
      }
      xmlChar *val = decode_text(vs, (size_t)(x->p - vs));
      x->p++;
      xmlAttrPtr a = (xmlAttrPtr)calloc(1, sizeof(xmlAttr));
      a->type = XML_ATTRIBUTE_NODE;
      a->name = an;
      a->parent = el;
      a->children = a->last = new_text(val);
      a->prev = last_attr;
      if (last_attr) last_attr->next = a;
      else el->properties = a;
      last_attr = a;
   }
   if (x->error) {
      free_node_list(el);
      return NULL;
   }

   parse_content(x, el);
   if (x->error) {
      free_node_list(el);
      return NULL;
   }
   /* expect </name> */
   if (!starts(x, "</")) {
      x->error = 1;
      free_node_list(el);
      return NULL;
   }
   x->p += 2;
   xmlChar *cn = parse_name(x);
   if (!cn || strcmp((char *)cn, (char *)el->name)) x->error = 1;
   free(cn);
   skip_ws(x);
   if (x->p < x->end && *x->p == '>') x->p++;
   else x->error = 1;
   if (x->error) {
      free_node_list(el);
      return NULL;
   }
   return el;
}

/* Parse children of `parent` until a closing tag or end of input. */
static void parse_content(xp *x, xmlNodePtr parent) {
   while (x->p < x->end && !x->error) {
      if (*x->p != '<') {
         const char *s = x->p;
         while (x->p < x->end && *x->p != '<') x->p++;
         xmlNodePtr t = new_text(decode_text(s, (size_t)(x->p - s)));
         link_child(parent, t);
         continue;
      }
      if (starts(x, "</")) return;
      if (starts(x, "<!--")) {
         const char *s = x->p + 4;
         x->p += 4;
         if (skip_until(x, "-->")) return;
         xmlNodePtr c = new_node(XML_COMMENT_NODE, dup_z((const xmlChar *)"comment"));
         c->content = dup_n(s, (size_t)(x->p - 3 - s));
         link_child(parent, c);
         continue;
      }
      if (starts(x, "<![CDATA[")) {
         const char *s = x->p + 9;
         x->p += 9;
         if (skip_until(x, "]]>")) return;
         link_child(parent, new_text(dup_n(s, (size_t)(x->p - 3 - s))));
         continue;
      }
      if (starts(x, "<?")) {
         if (skip_until(x, "?>")) return;
         continue;
      }
      if (starts(x, "<!")) { /* DOCTYPE etc.: skip to '>' */
         if (skip_until(x, ">")) return;
         continue;
      }
      xmlNodePtr el = parse_element(x);
      if (!el) {
         x->error = 1;
         return;
      }
      link_child(parent, el);
   }
}

xmlDocPtr xmlReadMemory(const char *buffer, int size, const char *URL, const char *encoding,
                        int options) {
   (void)encoding;
   (void)options;
   if (!buffer || size < 0) return NULL;
   xmlDocPtr doc = (xmlDocPtr)calloc(1, sizeof(xmlDoc));
   if (!doc) return NULL;
   doc->type = XML_DOCUMENT_NODE;
   doc->name = URL ? (char *)dup_z((const xmlChar *)URL) : NULL;

   /* Parse top level into a scratch element, then move children onto the doc. */
   xmlNode top;
   memset(&top, 0, sizeof(top));
   top.type = XML_ELEMENT_NODE;
   xp x = {buffer, buffer + size, 0};
   parse_content(&x, &top);
   if (!x.error && x.p < x.end) x.error = 1; /* stray closing tag */

   int nroot = 0;
   for (xmlNodePtr n = top.children; n; n = n->next)
      if (n->type == XML_ELEMENT_NODE) nroot++;
   if (x.error || nroot != 1) {
      fprintf(stderr, "%s: XML parse error near byte %ld\n", URL ? URL : "(memory)",
              (long)(x.p - buffer));
      free_node_list(top.children);
      free(doc->name);
      free(doc);
      return NULL;
   }
   doc->children = top.children;
   doc->last = top.last;
   for (xmlNodePtr n = doc->children; n; n = n->next) n->parent = NULL;
   set_doc_rec(doc->children, doc);
   return doc;
}

xmlNodePtr xmlDocGetRootElement(xmlDocPtr doc) {
   if (!doc) return NULL;
   for (xmlNodePtr n = doc->children; n; n = n->next)
      if (n->type == XML_ELEMENT_NODE) return n;
   return NULL;
}

xmlNodePtr xmlDocSetRootElement(xmlDocPtr doc, xmlNodePtr root) {
   if (!doc || !root) return NULL;
   xmlNodePtr old = xmlDocGetRootElement(doc);
   if (old) {
      /* unlink old root */
      if (old->prev) old->prev->next = old->next;
      else doc->children = old->next;
      if (old->next) old->next->prev = old->prev;
      else doc->last = old->prev;
      old->next = old->prev = NULL;
   }
   root->parent = NULL;
   root->next = NULL;
   root->prev = doc->last;
   if (doc->last) doc->last->next = root;
   else doc->children = root;
   doc->last = root;
   xmlNodePtr sib = root->next;
   root->next = NULL;
   set_doc_rec(root, doc);
   root->next = sib;
   return old;
}

xmlDocPtr xmlNewDoc(const xmlChar *version) {
   xmlDocPtr doc = (xmlDocPtr)calloc(1, sizeof(xmlDoc));
   if (!doc) return NULL;
   doc->type = XML_DOCUMENT_NODE;
   doc->version = dup_z(version ? version : (const xmlChar *)"1.0");
   return doc;
}

void xmlFreeDoc(xmlDocPtr doc) {
   if (!doc) return;
   free_node_list(doc->children);
   free(doc->name);
   free(doc->version);
   free(doc);
}

void xmlCleanupParser(void) {}

xmlChar *xmlGetProp(xmlNodePtr node, const xmlChar *name) {
   if (!node || !name) return NULL;
   for (xmlAttrPtr a = node->properties; a; a = a->next)
      if (!strcmp((const char *)a->name, (const char *)name))
         return dup_z(a->children && a->children->content ? a->children->content
                                                          : (const xmlChar *)"");
   return NULL;
}

xmlAttrPtr xmlNewProp(xmlNodePtr node, const xmlChar *name, const xmlChar *value) {
   if (!node || !name) return NULL;
   xmlAttrPtr a = (xmlAttrPtr)calloc(1, sizeof(xmlAttr));
   if (!a) return NULL;
   a->type = XML_ATTRIBUTE_NODE;
   a->name = dup_z(name);
   a->parent = node;
   a->doc = node->doc;
   a->children = a->last = new_text(dup_z(value ? value : (const xmlChar *)""));
   xmlAttrPtr t = node->properties;
   if (!t) node->properties = a;
   else {
      while (t->next) t = t->next;
      t->next = a;
      a->prev = t;
   }
   return a;
}

static void content_rec(xmlNodePtr n, char **buf, size_t *len, size_t *cap) {
   for (; n; n = n->next) {
      if (n->type == XML_TEXT_NODE && n->content) {
         size_t l = strlen((char *)n->content);
         if (*len + l + 1 > *cap) {
            *cap = (*len + l + 1) * 2;
            *buf = (char *)realloc(*buf, *cap);
         }
         memcpy(*buf + *len, n->content, l);
         *len += l;
         (*buf)[*len] = 0;
      } else if (n->type == XML_ELEMENT_NODE) {
         content_rec(n->children, buf, len, cap);
      }
   }
}

xmlChar *xmlNodeGetContent(xmlNodePtr cur) {
   if (!cur) return NULL;
   if (cur->type == XML_TEXT_NODE || cur->type == XML_COMMENT_NODE)
      return dup_z(cur->content ? cur->content : (const xmlChar *)"");
   size_t len = 0, cap = 64;
   char *buf = (char *)malloc(cap);
   buf[0] = 0;
   content_rec(cur->children, &buf, &len, &cap);
   return (xmlChar *)buf;
}

int xmlIsBlankNode(xmlNodePtr node) {
   if (!node || node->type != XML_TEXT_NODE) return 0;
   if (!node->content) return 1;
   for (const xmlChar *c = node->content; *c; c++)
      if (!isspace(*c)) return 0;
   return 1;
}

xmlNodePtr xmlNewNode(xmlNsPtr ns, const xmlChar *name) {
   (void)ns;
   if (!name) return NULL;
   return new_node(XML_ELEMENT_NODE, dup_z(name));
}

xmlNodePtr xmlAddChild(xmlNodePtr parent, xmlNodePtr cur) {
   if (!parent || !cur) return NULL;
   link_child(parent, cur);
   xmlNodePtr sib = cur->next;
   cur->next = NULL;
   set_doc_rec(cur, parent->doc);
   cur->next = sib;
   return cur;
}

xmlNodePtr xmlNewChild(xmlNodePtr parent, xmlNsPtr ns, const xmlChar *name,
                       const xmlChar *content) {
   if (!parent) return NULL;
   xmlNodePtr n = xmlNewNode(ns, name);
   if (!n) return NULL;
   if (content) link_child(n, new_text(dup_z(content)));
   return xmlAddChild(parent, n);
}

xmlNodePtr xmlCopyNode(xmlNodePtr node, int recursive) {
   if (!node) return NULL;
   xmlNodePtr c = new_node(node->type, dup_z(node->name));
   if (!c) return NULL;
   c->content = dup_z(node->content);
   for (xmlAttrPtr a = node->properties; a; a = a->next)
      xmlNewProp(c, a->name, a->children ? a->children->content : NULL);
   if (recursive)
      for (xmlNodePtr k = node->children; k; k = k->next) {
         xmlNodePtr kc = xmlCopyNode(k, recursive);
         if (kc) link_child(c, kc);
      }
   return c;
}

int xmlStrcmp(const xmlChar *a, const xmlChar *b) {
   if (a == b) return 0;
   if (!a) return -1;
   if (!b) return 1;
   return strcmp((const char *)a, (const char *)b);
}
